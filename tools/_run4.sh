set -x
timeout 400 python tools/sweep.py 28 keys=0,11,12,13,14,15,16,17,18 pairs=0,8,9,10,11,12,13,14,15,16 scan= > gpurun_out/s6_sweep.log 2>&1
B200RS_KEYS_VARIANT=11 B200RS_PAIRS_VARIANT=8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s6_pytest_a.log
B200RS_KEYS_VARIANT=13 B200RS_PAIRS_VARIANT=10 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s6_pytest_b.log
