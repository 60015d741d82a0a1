set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s7_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'onesweep' -s 9 -c 2 -o gpurun_out/s7_v2 python tools/prof_once.py 26 keys,pairs > gpurun_out/s7_ncu.log 2>&1
