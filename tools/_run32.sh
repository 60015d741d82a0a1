mkdir -p gpurun_out
timeout 600 python tools/sweep.py 28 dist=uniform keys=38,25,32 pairs=21,8,19 scan= > gpurun_out/s35_sweep.txt 2>&1
timeout 600 python tools/sweep.py 28 dist=sorted keys=38 pairs=21 scan= >> gpurun_out/s35_sweep.txt 2>&1
(time timeout 1200 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6 > gpurun_out/s35_pytest.log
