mkdir -p gpurun_out
timeout 300 tools/_build/dist_dropin 8 500003 | tail -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2k_bench_8gpu.json 2> gpurun_out/r2k_bench_8gpu.err; tail -3 gpurun_out/r2k_bench_8gpu.err; cat gpurun_out/r2k_bench_8gpu.json
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
for ov in 1 0; do B200RS_DIST_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$ov tools/dist_perf.py 28 p2p/dest 6 2>&1 | grep GPUs | sed "s/^/overlap=$ov /" >> gpurun_out/r2k_dist_8gpu_phases.txt; done
B200RS_XP_NO_BULK=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29615 tools/dist_perf.py 28 p2p/dest 6 2>&1 | grep -E "GPUs|local sort" | sed "s/^/nobulk /" >> gpurun_out/r2k_dist_8gpu_phases.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29616 tools/dist_perf.py 31 p2p/dest 4 2>&1 | grep -E "GPUs|local sort" >> gpurun_out/r2k_dist_8gpu_phases.txt
cat gpurun_out/r2k_dist_8gpu_phases.txt
