mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2h_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2h_pytest_gpu.log
tools/_build/prims_dropin | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -3 gpurun_out/r2h_bench.err; cat gpurun_out/r2h_bench.json
timeout 600 python tools/msd_probe.py perf 28 uniform sorted > gpurun_out/r2h_msd_perf.txt 2>&1; cat gpurun_out/r2h_msd_perf.txt
