mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_dist.py -x -q) > gpurun_out/r2j_pytest_dist.log 2>&1; tail -5 gpurun_out/r2j_pytest_dist.log
timeout 300 tools/_build/dist_dropin 2 1000003 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-config5 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; tail -3 gpurun_out/r2j_bench_2gpu.err; cut -c1-400 gpurun_out/r2j_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/dist_perf.py 28 p2p/dest 6 > gpurun_out/r2j_dist_2gpu_phases.txt 2>&1; tail -20 gpurun_out/r2j_dist_2gpu_phases.txt
