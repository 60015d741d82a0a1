mkdir -p gpurun_out
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
timeout 600 python tools/msd_probe.py check 2>&1 | tail -3
for h in 0 1 2 3 4 5; do echo "H shape $h"; B200RS_MSD_H=$h timeout 300 python tools/msd_probe.py perf 28 uniform 2>&1 | tail -1; done
echo "no PDL"; B200RS_MSD_NO_PDL=1 timeout 300 python tools/msd_probe.py perf 28 uniform sorted 2>&1 | tail -2
echo "PDL"; timeout 300 python tools/msd_probe.py perf 28 uniform sorted and3 2>&1 | tail -3
timeout 300 python tools/msd_probe.py perf 27 29 uniform 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'msd_hist16' -s 1 -c 1 -o /tmp/r2n_h python tools/msd_prof_once.py 28 uniform > gpurun_out/r2n_ncu_h.log 2>&1
ncu -i /tmp/r2n_h.ncu-rep --page raw --csv > gpurun_out/r2n_raw_h.csv 2>/dev/null
ncu -i /tmp/r2n_h.ncu-rep --page source --csv > gpurun_out/r2n_source_h.csv 2>/dev/null
ls -la gpurun_out/r2n_*
