# usage: bash tools/scripts/bench_nN.sh N [workloads...]   (train config3 config5 config5-accumulated)
N=$1; shift
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for w in "$@"; do
  case $w in
    train) $T bench.py --gpus $N --mode train --steps 10 --warmup 3 > gpurun_out/bench_r2_train_n$N.json 2> gpurun_out/bench_r2_train_n$N.err;;
    config3) $T bench.py --gpus $N --no-cpu-baseline --no-hbm > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err;;
    *) $T bench.py --gpus $N --workload $w --no-cpu-baseline --no-hbm > gpurun_out/bench_r2_${w}_n$N.json 2> gpurun_out/bench_r2_${w}_n$N.err;;
  esac
  echo "$w rc=$?"
done
for f in gpurun_out/bench_r2_*_n$N.json gpurun_out/bench_r2_n$N.json; do [ -f $f ] && (echo $f; tail -1 $f | cut -c1-200); done
