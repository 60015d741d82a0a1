mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
export B200RS_LIB=$PWD/tools/_build/libb200rs_exp.so
for v in 0 1 2; do echo "pairs regular-pass variant $v"; B200RS_PAIRS_REGULAR=$v timeout 300 python tools/pair_distributions.py 28 uniform sorted reversed and3 2>&1 | tail -4; done
