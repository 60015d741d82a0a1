mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/dist_perf.py 28 p2p/dest 6 2>&1 | grep -E "GPUs|local sort|rror" | sed "s/^/$2 /" >> gpurun_out/r2o_dist_2gpu_phases.txt; }
run 29621 pipelined
B200RS_DIST_SEQUENTIAL_SORTS=1 run 29622 pipelined-sequential-sorts
B200RS_DIST_NO_PIPELINE=1 run 29623 unpipelined
cat gpurun_out/r2o_dist_2gpu_phases.txt
unset B200RS_LIB
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus 2 --steps 10 --warmup 3 --no-config5 > gpurun_out/r2o_bench_2gpu.json 2> gpurun_out/r2o_bench_2gpu.err; tail -3 gpurun_out/r2o_bench_2gpu.err; cut -c1-1500 gpurun_out/r2o_bench_2gpu.json
