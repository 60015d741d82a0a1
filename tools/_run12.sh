timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s14_pytest.log
timeout 300 python tools/sweep.py 28 keys=0,14 pairs=0,8 scan= > gpurun_out/s14_sweep.log 2>&1
