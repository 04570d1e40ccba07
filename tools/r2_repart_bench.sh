#!/usr/bin/env bash
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for extra in "" "--no-presort"; do
  for env in "BH_REORDER=1" "BH_REORDER=0"; do
    echo "== $extra $env"
    env $env timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --quick --total-bodies 4000000 $extra 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step %.4f' % d['ms_per_step'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], d['phases_us'])"
  done
done 2>&1 | tee gpurun_out/r2_repart_bench_g$N.log
