mkdir -p gpurun_out
timeout 600 python tools/sweep.py 28 dist=sorted keys=25,38,40,41,42 pairs=8,21,30,31 scan= > gpurun_out/s30_sweep.txt 2>&1
timeout 600 python tools/sweep.py 28 dist=uniform keys=25,38,40,41,42 pairs=8,21,30,31 scan= >> gpurun_out/s30_sweep.txt 2>&1
B200RS_KEYS_VARIANT=41 B200RS_PAIRS_VARIANT=31 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/s30_pytest_swz.log
