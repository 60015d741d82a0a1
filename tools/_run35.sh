mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/s37_ngpus.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s37_bench8.json 2> gpurun_out/s37_bench8.err
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/dist_perf.py 28 p2p/dest 6 > gpurun_out/s37_dist_perf8_28.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/dist_perf.py 31 p2p/dest 4 > gpurun_out/s37_dist_perf8_31.log 2>&1
