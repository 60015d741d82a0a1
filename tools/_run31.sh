mkdir -p gpurun_out
for s in 20 22 24 26; do timeout 300 python tools/quick_perf.py $s > gpurun_out/s34_quick_$s.txt 2>&1; done
