mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2e_pytest_gpu.log
timeout 600 python tools/msd_probe.py perf 28 uniform sorted and3 > gpurun_out/r2e_msd_perf.txt 2>&1
timeout 600 python tools/msd_probe.py perf 27 29 uniform >> gpurun_out/r2e_msd_perf.txt 2>&1
cat gpurun_out/r2e_msd_perf.txt
timeout 600 python tools/key_distributions.py 28 > gpurun_out/r2e_key_distributions.md 2>&1; cat gpurun_out/r2e_key_distributions.md
