mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/host_sweep.py > gpurun_out/r2l_host_sweep_coop.md 2>&1; tail -24 gpurun_out/r2l_host_sweep_coop.md
B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so B200RS_NO_COOP_MID=1 timeout 300 python tools/host_sweep.py > gpurun_out/r2l_host_sweep_chain.md 2>&1; tail -24 gpurun_out/r2l_host_sweep_chain.md | cut -d'|' -f2,3,6,8,10
timeout 300 python tools/msd_probe.py perf 28 uniform sorted 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -2 gpurun_out/r2l_bench.err; cat gpurun_out/r2l_bench.json
