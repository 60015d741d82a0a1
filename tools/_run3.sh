set -x
timeout 300 python tools/sweep.py 28 keys=0,11,12 pairs=0,8,9,11 scan= > gpurun_out/s5_sweep.log 2>&1
B200RS_KEYS_VARIANT=11 B200RS_PAIRS_VARIANT=8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s5_pytest_v2.log
B200RS_KEYS_VARIANT=11 B200RS_PAIRS_VARIANT=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'onesweep' -s 9 -c 2 -o gpurun_out/s5_v2 python tools/prof_once.py 26 keys,pairs > gpurun_out/s5_ncu.log 2>&1
