mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6 > gpurun_out/s31_pytest.log
B200RS_TOOL_PROFILE=1 timeout 900 python tools/key_distributions.py 28 30 > gpurun_out/s31_dist.txt 2>&1
timeout 600 python tools/big_sizes.py > gpurun_out/s31_big.txt 2>&1
