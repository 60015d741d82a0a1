timeout 600 python tools/sweep.py 28 keys= pairs= scan=0,12,14,17,19,20,21,22,23,24,25,26 > gpurun_out/s10_sweep.log 2>&1
B200RS_SCAN_VARIANT=19 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k scan 2>&1 | tail -5 > gpurun_out/s10_pytest.log
B200RS_SCAN_VARIANT=24 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k scan 2>&1 | tail -5 >> gpurun_out/s10_pytest.log
