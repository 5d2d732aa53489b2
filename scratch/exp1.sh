set -x
python tools/bench_conv.py > gpurun_out/bench_conv_d1.log 2>&1; cat gpurun_out/bench_conv_d1.log
LTN_CONV_DEPTH=2 python tools/bench_conv.py > gpurun_out/bench_conv_d2.log 2>&1; cat gpurun_out/bench_conv_d2.log
python -m pytest tests/test_conv_tc_gpu.py -x -q 2>&1 | tail -2
LTN_CONV_DEPTH=2 python -m pytest tests/test_conv_tc_gpu.py -x -q 2>&1 | tail -2
for L in 1 2 3 4 6; do python bench.py --steps 24 --warmup 3 --no-cpu-baseline --lanes $L 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lanes',$L,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
"; done
