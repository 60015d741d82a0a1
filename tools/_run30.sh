mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/s33_ngpus.log
(time timeout 600 python -m pytest tests/test_gpu_dist.py -x -q) 2>&1 | tail -6 > gpurun_out/s33_pytest_dist.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s33_bench2.json 2> gpurun_out/s33_bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/dist_perf.py 28 > gpurun_out/s33_dist_perf2.log 2>&1
