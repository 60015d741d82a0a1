mkdir -p gpurun_out
for pf in 0 148 296 592 888 1184 2368; do
  echo "== B200RS_PF_TILES=$pf" >> gpurun_out/s24_sweep.txt
  B200RS_PF_TILES=$pf timeout 300 python tools/sweep.py 28 keys=25 pairs=8,19 scan= >> gpurun_out/s24_sweep.txt 2>&1
done
echo "== default" >> gpurun_out/s24_sweep.txt
timeout 300 python tools/sweep.py 28 keys=25,28,22 pairs=8,14,17 scan= >> gpurun_out/s24_sweep.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/s24_pytest.log
