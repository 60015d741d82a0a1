mkdir -p gpurun_out
timeout 600 python tools/size_scaling.py > gpurun_out/s32_sizes.txt 2>&1
