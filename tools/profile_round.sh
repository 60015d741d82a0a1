# tools/profile_round.sh -- what produced profiles/r1c_*: GPU tests, smoke, bench (both arms), the ncu launch list and the ncu --set full
# captures (reports stay in /tmp on the GPU box: only the raw / source CSV pages are brought back).  Run: gpurun -- bash tools/profile_round.sh
mkdir -p gpurun_out
set -x
(time timeout 1200 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6 > gpurun_out/r1c_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1c_smoke.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1c_bench_ref.json 2>> gpurun_out/r1c_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r1c_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'onesweep2|digit_histogram' -s 5 -c 2 -o /tmp/r1c_full_pairs python tools/prof_once.py 28 pairs > gpurun_out/r1c_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'onesweep2|digit_histogram' -s 5 -c 2 -o /tmp/r1c_full_keys python tools/prof_once.py 28 keys >> gpurun_out/r1c_ncu.log 2>&1
ncu --set full --clock-control none -k regex:'scan_ring|copy_u32|fill_kernel' -c 4 -o /tmp/r1c_full_misc python tools/prof_misc.py >> gpurun_out/r1c_ncu.log 2>&1
for r in pairs keys misc; do ncu -i /tmp/r1c_full_$r.ncu-rep --page raw --csv > gpurun_out/r1c_raw_$r.csv 2>/dev/null; done
ncu -i /tmp/r1c_full_pairs.ncu-rep --page source --csv --kernel-name regex:onesweep2 --launch-count 1 > gpurun_out/r1c_source_pairs.csv 2>/dev/null
ncu -i /tmp/r1c_full_keys.ncu-rep --page source --csv --kernel-name regex:onesweep2 --launch-count 1 > gpurun_out/r1c_source_keys.csv 2>/dev/null
ls -la /tmp/*.ncu-rep; du -sh gpurun_out
