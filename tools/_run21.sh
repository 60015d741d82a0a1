mkdir -p gpurun_out
for pf in 16 32 64 96 148 222; do
  echo "== B200RS_PF_TILES=$pf" >> gpurun_out/s25_sweep.txt
  B200RS_PF_TILES=$pf timeout 300 python tools/sweep.py 28 keys=25 pairs=8 scan= >> gpurun_out/s25_sweep.txt 2>&1
done
for pf in 64 148; do
  echo "== shapes, B200RS_PF_TILES=$pf" >> gpurun_out/s25_sweep.txt
  B200RS_PF_TILES=$pf timeout 600 python tools/sweep.py 28 keys=25,22,28,29,30,32,14,16,18,23,26 pairs=8,11,14,17,18,19,21,22 scan= >> gpurun_out/s25_sweep.txt 2>&1
done
