mkdir -p gpurun_out
timeout 300 tools/_build/dist_dropin 8 3400007 | tail -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2t_bench_8gpu.json 2> gpurun_out/r2t_bench_8gpu.err; tail -3 gpurun_out/r2t_bench_8gpu.err; cat gpurun_out/r2t_bench_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tools/dist_perf.py 28 p2p/dest 6 2>&1 | grep -E "GPUs|local sort" >> gpurun_out/r2t_dist_8gpu_phases.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29616 tools/dist_perf.py 31 p2p/dest 4 2>&1 | grep -E "GPUs|local sort" >> gpurun_out/r2t_dist_8gpu_phases.txt
cat gpurun_out/r2t_dist_8gpu_phases.txt
