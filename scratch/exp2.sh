for E in 0 1 2 3; do for D in 1 2; do echo "== EXP $E DEPTH $D"; LTN_CONV_EXP=$E LTN_CONV_DEPTH=$D python tools/bench_conv.py 2>&1 | cut -c1-90; done; done
for E in 2 3; do LTN_CONV_EXP=$E python -m pytest tests/test_conv_tc_gpu.py -x -q 2>&1 | tail -2; done
