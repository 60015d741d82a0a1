mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/s21_gpu.txt
(time timeout 1200 python -m pytest tests -m gpu -x -q) 2>&1 | tail -8 > gpurun_out/s21_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err
timeout 300 tools/_build/cub_compare 28 10 > gpurun_out/s21_cub_28.json 2> gpurun_out/s21_cub.err
timeout 300 tools/_build/cub_compare 30 5 > gpurun_out/s21_cub_30.json 2>> gpurun_out/s21_cub.err
timeout 300 python tools/quick_perf.py 28 > gpurun_out/s21_quick_28.txt 2>&1
