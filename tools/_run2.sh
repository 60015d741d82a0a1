set -x
timeout 300 python tools/sweep.py 28 keys=0,11,12,13,14,15,16 pairs=0,8,9,10,11,12,13 scan= > gpurun_out/s4_sweep.log 2>&1
B200RS_KEYS_VARIANT=11 B200RS_PAIRS_VARIANT=8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s4_pytest_v2.log
