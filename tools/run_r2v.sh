mkdir -p gpurun_out
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
for a in 500 440 400; do B200RS_DIST_A_PERMILLE=$a timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961${a:1:1} tools/dist_perf.py 28 p2p/dest 6 2>&1 | grep -E "GPUs" | sed "s/^/A=$a /" >> gpurun_out/r2v_dist_8gpu_split.txt; done
for a in 500 440; do B200RS_DIST_A_PERMILLE=$a timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2962${a:1:1} tools/dist_perf.py 31 p2p/dest 4 2>&1 | grep -E "GPUs" | sed "s/^/A=$a /" >> gpurun_out/r2v_dist_8gpu_split.txt; done
cat gpurun_out/r2v_dist_8gpu_split.txt
