nvidia-smi -L > gpurun_out/s13_gpus.log
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s13_pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_perf.py 28 > gpurun_out/s13_dist_perf.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s13_bench2.log 2>&1
