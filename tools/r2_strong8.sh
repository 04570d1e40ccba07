#!/usr/bin/env bash
set -u
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --quick --total-bodies 16000000 > gpurun_out/r2_strong16M_g$N.json 2> gpurun_out/r2_strong16M_g$N.err
tail -2 gpurun_out/r2_strong16M_g$N.err; grep '^{' gpurun_out/r2_strong16M_g$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', d['ms_per_step'], 'value', d['value'])
for r,p in enumerate(d['phases_us_all_ranks']): print(r, p)
"
