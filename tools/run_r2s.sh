mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -5
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/dist_perf.py 28 p2p/dest 6 2>&1 | grep -E "GPUs|local sort|rror" | sed "s/^/$2 /" >> gpurun_out/r2s_dist_2gpu_phases.txt; }
run 29621 pipelined
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
B200RS_DIST_NO_PIPELINE=1 run 29623 unpipelined
unset B200RS_LIB
cat gpurun_out/r2s_dist_2gpu_phases.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus 2 --steps 10 --warmup 3 --no-config5 > gpurun_out/r2s_bench_2gpu.json 2> gpurun_out/r2s_bench_2gpu.err; tail -3 gpurun_out/r2s_bench_2gpu.err; cut -c1-400 gpurun_out/r2s_bench_2gpu.json
