mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_msd.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2w_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2w_bench_2gpu.json 2> gpurun_out/r2w_bench_2gpu.err; tail -2 gpurun_out/r2w_bench_2gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r2w_bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['parity'], d.get('config5'))"
