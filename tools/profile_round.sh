# tools/profile_round.sh -- what produced profiles/r2_final_*: GPU tests, smoke, bench (both arms), the ncu launch list of the bench
# command and the ncu --set full captures (reports stay in /tmp on the GPU box: only the raw / source CSV pages are brought back).
# Run: gpurun -- bash tools/profile_round.sh     (one GPU, ~6 minutes)
mkdir -p gpurun_out
P=gpurun_out/r2_final
(time timeout 1500 python -m pytest tests -m gpu -q) 2>&1 | tail -6 > ${P}_pytest_gpu.log; cat ${P}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -2 ${P}_smoke.log
python bench.py --steps 10 --warmup 3 > ${P}_bench.json 2> ${P}_bench.err; tail -2 ${P}_bench.err; cut -c1-300 ${P}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference_arm.json 2>> ${P}_bench.err; cut -c1-300 ${P}_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > ${P}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'onesweep2|digit_histogram' -s 5 -c 2 -o /tmp/r2_full_pairs python tools/prof_once.py 28 pairs > ${P}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'msd_' -s 6 -c 6 -o /tmp/r2_full_keys python tools/msd_prof_once.py 28 uniform >> ${P}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'exchange_partition' -s 1 -c 1 -o /tmp/r2_full_xp python tools/xp_probe.py 28 16 once >> ${P}_ncu.log 2>&1
ncu --set full --clock-control none -k regex:'scan_ring|copy_u32|fill_kernel' -c 4 -o /tmp/r2_full_misc python tools/prof_misc.py >> ${P}_ncu.log 2>&1
for r in pairs keys xp misc; do ncu -i /tmp/r2_full_$r.ncu-rep --page raw --csv > ${P}_raw_$r.csv 2>/dev/null; done
ncu -i /tmp/r2_full_xp.ncu-rep --page source --csv > ${P}_source_xp.csv 2>/dev/null
python tools/key_distributions.py 28 > ${P}_key_distributions.md 2>&1; cat ${P}_key_distributions.md
python tools/pair_distributions.py 28 > ${P}_pair_distributions.md 2>&1; cat ${P}_pair_distributions.md
python tools/host_sweep.py > ${P}_config0_host_sweep.md 2>&1
python tools/xp_probe.py 28 2 8 16 32 > ${P}_exchange_local.txt 2>&1; cat ${P}_exchange_local.txt
du -sh gpurun_out
