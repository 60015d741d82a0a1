mkdir -p gpurun_out
timeout 600 python tools/msd_probe.py check > gpurun_out/r2g_msd_check.txt 2>&1; echo "check rc=$?"; tail -1 gpurun_out/r2g_msd_check.txt
timeout 600 python tools/msd_probe.py perf 28 uniform sorted and3 allequal > gpurun_out/r2g_msd_perf.txt 2>&1
timeout 600 python tools/msd_probe.py perf 27 29 uniform >> gpurun_out/r2g_msd_perf.txt 2>&1
cat gpurun_out/r2g_msd_perf.txt
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
for pf in 0 148 296 592 1184; do B200RS_MSD_PPF=$pf timeout 300 python tools/msd_probe.py perf 28 uniform 2>&1 | sed "s/^/PPF$pf /" >> gpurun_out/r2g_msd_shapes.txt; done
for p in 10 11; do B200RS_MSD_P=$p timeout 300 python tools/msd_probe.py perf 28 uniform 2>&1 | sed "s/^/P$p /" >> gpurun_out/r2g_msd_shapes.txt; done
cat gpurun_out/r2g_msd_shapes.txt
