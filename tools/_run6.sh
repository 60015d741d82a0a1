set -x
timeout 600 python tools/sweep.py 28 keys=0,11,12,14,15,16,17,18,19,20,21 pairs=0,8,9,11,12,13,14,15,16 scan= > gpurun_out/s8_sweep.log 2>&1
B200RS_KEYS_VARIANT=15 B200RS_PAIRS_VARIANT=9 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s8_pytest.log
