mkdir -p gpurun_out
set -x
timeout 600 python tools/msd_probe.py check > gpurun_out/r2b_msd_check.txt 2>&1; echo "check rc=$?"; tail -3 gpurun_out/r2b_msd_check.txt
timeout 600 python tools/msd_probe.py perf 28 uniform sorted and3 > gpurun_out/r2b_msd_perf.txt 2>&1
timeout 600 python tools/msd_probe.py perf 26 27 29 uniform >> gpurun_out/r2b_msd_perf.txt 2>&1
cat gpurun_out/r2b_msd_perf.txt
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
for p in 7 8 9 10 11 12; do B200RS_MSD_P=$p timeout 300 python tools/msd_probe.py perf 28 uniform 2>&1 | sed "s/^/P$p /" >> gpurun_out/r2b_msd_shapes.txt; done
for f in 0 1 4 5 6 7 9; do B200RS_MSD_F=$f timeout 300 python tools/msd_probe.py perf 28 uniform 2>&1 | sed "s/^/F$f /" >> gpurun_out/r2b_msd_shapes.txt; done
for pf in 0 148 444 2048; do B200RS_MSD_PF=$pf timeout 300 python tools/msd_probe.py perf 28 uniform 2>&1 | sed "s/^/PF$pf /" >> gpurun_out/r2b_msd_shapes.txt; done
cat gpurun_out/r2b_msd_shapes.txt
unset B200RS_LIB
for kind in uniform sorted; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'msd_partition|msd_bucket' -s 3 -c 3 -o /tmp/r2b_$kind python tools/msd_prof_once.py 28 $kind > gpurun_out/r2b_ncu_$kind.log 2>&1
  ncu -i /tmp/r2b_$kind.ncu-rep --page raw --csv > gpurun_out/r2b_raw_$kind.csv 2>/dev/null
  ncu -i /tmp/r2b_$kind.ncu-rep --page source --csv --kernel-name regex:msd_partition --launch-count 1 > gpurun_out/r2b_source_partition_$kind.csv 2>/dev/null
  ncu -i /tmp/r2b_$kind.ncu-rep --page source --csv --kernel-name regex:msd_bucket --launch-count 1 > gpurun_out/r2b_source_bucket_$kind.csv 2>/dev/null
done
ls -la /tmp/*.ncu-rep; du -sh gpurun_out
