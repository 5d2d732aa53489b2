T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --mode train --steps 10 --warmup 3 > gpurun_out/bench_r2_train_n2.json 2> gpurun_out/bench_r2_train_n2.err; echo train rc=$?
$T bench.py --gpus 2 --no-cpu-baseline --no-hbm > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo c3 rc=$?
$T bench.py --gpus 2 --workload config5 --no-cpu-baseline --no-hbm > gpurun_out/bench_r2_config5_n2.json 2> gpurun_out/bench_r2_config5_n2.err; echo c5 rc=$?
$T bench.py --gpus 2 --workload config5-accumulated --no-cpu-baseline --no-hbm > gpurun_out/bench_r2_config5-accumulated_n2.json 2> gpurun_out/bench_r2_config5acc_n2.err; echo c5a rc=$?
for f in gpurun_out/bench_r2_train_n2.json gpurun_out/bench_r2_n2.json gpurun_out/bench_r2_config5_n2.json gpurun_out/bench_r2_config5-accumulated_n2.json; do tail -1 $f | cut -c1-230; done
