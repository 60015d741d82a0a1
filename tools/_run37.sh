mkdir -p gpurun_out
for st in 1 0; do
echo "== B200RS_PASS0_STABLE=$st" >> gpurun_out/s39_unord.txt
B200RS_PASS0_STABLE=$st B200RS_TOOL_PROFILE=1 timeout 600 python tools/key_distributions.py 28 >> gpurun_out/s39_unord.txt 2>&1
done
(time timeout 1200 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6 >> gpurun_out/s39_unord.txt
