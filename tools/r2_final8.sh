#!/usr/bin/env bash
set -u
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
out=gpurun_out/r2_multi_gpu_check_g$N.log
: > $out
for args in "--bodies 200000 --steps 3" "--bodies 1500000 --steps 2 --host-step"; do
  echo "== multi_gpu_check $args" >> $out
  timeout 300 $TR tests/multi_gpu_check.py $args 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version" | tail -10 >> $out
done
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_g$N.json 2> gpurun_out/r2_bench_g$N.err
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --quick --no-e2e --total-bodies 64000000 > gpurun_out/r2_strong_64000000_g$N.json 2> gpurun_out/r2_strong_64000000_g$N.err
cat $out; tail -2 gpurun_out/r2_bench_g$N.err; grep '^{' gpurun_out/r2_bench_g$N.json | cut -c1-200; grep '^{' gpurun_out/r2_strong_64000000_g$N.json | cut -c1-200; tail -2 gpurun_out/r2_strong_64000000_g$N.err
