mkdir -p gpurun_out
timeout 600 python tools/sweep.py 28 dist=sorted keys=25,36,37,38,39,32,30 pairs=8,21,25,26,27,28,29 scan= > gpurun_out/s29_sweep.txt 2>&1
timeout 600 python tools/sweep.py 28 dist=uniform keys=25,36,37,38,39 pairs=8,21,25,26,27,28,29 scan= >> gpurun_out/s29_sweep.txt 2>&1
