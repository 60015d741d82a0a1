timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s12_pytest.log
for l in 16 20 22 24 26 28 30; do timeout 300 python tools/sweep.py $l keys= pairs= scan=0,20,24 ; done > gpurun_out/s12_scan_sizes.log 2>&1
