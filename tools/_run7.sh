timeout 600 python tools/sweep.py 28 keys=0,14,22,20,23 pairs=0,8,17,13,18 scan= > gpurun_out/s9_sweep.log 2>&1
