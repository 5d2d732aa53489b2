for i in 1 2 3 4 5 6; do python -m pytest tests/test_conv_tc_gpu.py tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/flaky_$i.log 2>&1; tail -1 gpurun_out/flaky_$i.log; done
grep -l failed gpurun_out/flaky_*.log | head -1 | xargs -r grep -E "^(E|FAILED|tests/)" | head -40
