mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
tools/_build/launch_cost > gpurun_out/r2a_launch_cost.txt 2>&1
timeout 600 python tools/msd_probe.py check > gpurun_out/r2a_msd_check.txt 2>&1; echo "check rc=$?"
tail -25 gpurun_out/r2a_msd_check.txt
timeout 600 python tools/msd_probe.py perf 28 uniform sorted reversed blocks and3 > gpurun_out/r2a_msd_perf.txt 2>&1
timeout 600 python tools/msd_probe.py perf 25 26 27 29 uniform >> gpurun_out/r2a_msd_perf.txt 2>&1
cat gpurun_out/r2a_msd_perf.txt
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
for p in 0 1 2 3 4 5 6 7; do B200RS_MSD_P=$p timeout 300 python tools/msd_probe.py perf 28 uniform 2>&1 | sed "s/^/P$p /" >> gpurun_out/r2a_msd_shapes.txt; done
for f in 0 3 4 5 1; do B200RS_MSD_F=$f timeout 300 python tools/msd_probe.py perf 28 uniform 2>&1 | sed "s/^/F$f /" >> gpurun_out/r2a_msd_shapes.txt; done
cat gpurun_out/r2a_msd_shapes.txt
unset B200RS_LIB
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2a_pytest_gpu.log
cat gpurun_out/r2a_launch_cost.txt
