mkdir -p gpurun_out
timeout 600 python tools/sweep.py 28 keys=25,36,37,38 pairs=8,25,26,27,28,29,30 scan= > gpurun_out/s26_sweep.txt 2>&1
