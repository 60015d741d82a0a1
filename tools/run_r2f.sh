mkdir -p gpurun_out
timeout 600 python tools/msd_probe.py check > gpurun_out/r2f_msd_check.txt 2>&1; echo "check rc=$?"; tail -1 gpurun_out/r2f_msd_check.txt; grep -c WRONG gpurun_out/r2f_msd_check.txt
timeout 900 python -m pytest tests/test_gpu_msd.py -x -q 2>&1 | tail -3
timeout 600 python tools/msd_probe.py perf 28 uniform sorted blocks staircase and3 allequal > gpurun_out/r2f_msd_perf.txt 2>&1
timeout 600 python tools/msd_probe.py perf 24 25 26 27 29 30 uniform >> gpurun_out/r2f_msd_perf.txt 2>&1
cat gpurun_out/r2f_msd_perf.txt
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
for f in 0 1 2 3 4 5 6 7; do B200RS_MSD_F=$f timeout 300 python tools/msd_probe.py perf 28 uniform 2>&1 | sed "s/^/F$f /" >> gpurun_out/r2f_msd_shapes.txt; done
for pf in 0 444 1776 7104; do B200RS_MSD_PF=$pf timeout 300 python tools/msd_probe.py perf 28 uniform 2>&1 | sed "s/^/PF$pf /" >> gpurun_out/r2f_msd_shapes.txt; done
cat gpurun_out/r2f_msd_shapes.txt
unset B200RS_LIB
for kind in uniform; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'msd_bucket|msd_hist' -s 2 -c 2 -o /tmp/r2f_$kind python tools/msd_prof_once.py 28 $kind > gpurun_out/r2f_ncu_$kind.log 2>&1
  ncu -i /tmp/r2f_$kind.ncu-rep --page raw --csv > gpurun_out/r2f_raw_$kind.csv 2>/dev/null
  ncu -i /tmp/r2f_$kind.ncu-rep --page source --csv --kernel-name regex:msd_bucket --launch-count 1 > gpurun_out/r2f_source_bucket_$kind.csv 2>/dev/null
done
