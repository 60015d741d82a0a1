mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29634 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2y_bench_4gpu.json 2> gpurun_out/r2y_bench_4gpu.err; tail -2 gpurun_out/r2y_bench_4gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r2y_bench_4gpu.json')); print(d['value'], d['ms_per_step'], d['parity'], d.get('config5'))"
