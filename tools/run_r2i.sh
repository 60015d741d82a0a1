mkdir -p gpurun_out
nvidia-smi -L | head -3
(time timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/r2i_pytest_gpu.log 2>&1; tail -15 gpurun_out/r2i_pytest_gpu.log
timeout 300 tools/_build/dist_dropin 2 1000003 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-config5 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err; tail -5 gpurun_out/r2i_bench_2gpu.err; cat gpurun_out/r2i_bench_2gpu.json | cut -c1-1500
timeout 600 python tools/msd_probe.py perf 28 uniform sorted > gpurun_out/r2i_msd_perf.txt 2>&1; cat gpurun_out/r2i_msd_perf.txt
