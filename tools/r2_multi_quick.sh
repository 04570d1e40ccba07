#!/usr/bin/env bash
set -u
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
out=gpurun_out/r2_multi_quick_g$N.log
: > $out
( timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -2 ) >> $out
for args in "--bodies 200000 --steps 3" "--bodies 200000 --steps 20 --fp64 --repartition" "--bodies 600000 --steps 70 --fp64 --repartition" "--bodies 600000 --steps 4 --repartition" "--bodies 1500000 --steps 2 --host-step" "--bodies 200000 --steps 2 --no-p2p" "--bodies 120000 --steps 3 --exact-leaves" "--bodies 120000 --steps 2 --exact-leaves --fp64"; do
  echo "== multi_gpu_check $args" >> $out
  timeout 300 $TR tests/multi_gpu_check.py $args 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version" | tail -5 >> $out
done
cat $out
bash tools/r2_strong8.sh $N
