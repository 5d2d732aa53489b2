for CL in 1 2 4; do echo "== CLUSTER $CL"; LTN_CONV_CLUSTER=$CL timeout 120 python tools/bench_conv.py 2>&1 | tail -7 | cut -c1-60; done
