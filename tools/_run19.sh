mkdir -p gpurun_out
timeout 300 python tools/sweep.py 28 keys=25,34,35,36,37,38,39 pairs=8,23,24,25,26,27,28 scan= > gpurun_out/s23_sweep.txt 2>&1
B200RS_KEYS_VARIANT=34 B200RS_PAIRS_VARIANT=23 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/s23_pytest_v3.log
