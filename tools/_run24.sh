mkdir -p gpurun_out
B200RS_TOOL_PROFILE=1 timeout 600 python tools/key_distributions.py 28 > gpurun_out/s27_dist.txt 2>&1
