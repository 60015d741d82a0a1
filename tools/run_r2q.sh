mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
for t in 128 256; do for pf in 0 64 128 256; do echo "threads $t pf $pf: $(B200RS_XP_THREADS=$t B200RS_XP_PF=$pf timeout 300 python tools/xp_probe.py 28 16 2>&1 | tail -1)"; done; done | tee gpurun_out/r2q_xp_shapes.txt
echo "threads 128 32 parts: $(timeout 300 python tools/xp_probe.py 28 32 2>&1 | tail -1)" | tee -a gpurun_out/r2q_xp_shapes.txt
