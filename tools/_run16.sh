timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s20_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/s20_bench.json 2> gpurun_out/s20_bench.err
