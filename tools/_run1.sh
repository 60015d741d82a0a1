set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s3_pytest.log
python tools/sweep.py 28 keys=0,3,4 pairs=0,1,3 scan=0,1,2,3,4,5,6,7,8,9,10,11 > gpurun_out/s3_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'onesweep|digit_histogram|scan_' -s 6 -c 6 -o gpurun_out/s3_full python tools/prof_once.py 26 > gpurun_out/s3_ncu.log 2>&1
