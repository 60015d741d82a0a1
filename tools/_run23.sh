mkdir -p gpurun_out
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r1c_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'onesweep2|digit_histogram|scan_ring' -s 10 -c 8 -o gpurun_out/r1c_full python tools/prof_once.py 28 > gpurun_out/r1c_ncu.log 2>&1
ls -la gpurun_out/
