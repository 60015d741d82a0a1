mkdir -p gpurun_out
for hv in 0 1; do
echo "== B200RS_HIST_VARIANT=$hv" >> gpurun_out/s36_hist.txt
B200RS_HIST_VARIANT=$hv timeout 300 python tools/quick_perf.py 28 2>&1 | grep -E "bits=|histogram|scan" >> gpurun_out/s36_hist.txt
B200RS_HIST_VARIANT=$hv timeout 300 python tools/quick_perf.py 24 2>&1 | grep -E "bits=|histogram" >> gpurun_out/s36_hist.txt
done
B200RS_HIST_VARIANT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 >> gpurun_out/s36_hist.txt
