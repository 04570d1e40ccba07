#!/usr/bin/env bash
# BASELINE config 3 in its physically meaningful form on N GPUs: Plummer 16M, depth cap 13, exact in-leaf pairs
# (every step: all-gather of the positions, full build on every rank, own-slice walk).  tools/r2_config3_g8.sh <ngpus>
set -u
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --brackets 5 --dist plummer --max-depth 13 --exact-leaves \
    --total-bodies 16000000 --no-p2p --no-e2e --no-strong --no-direct --no-gpu-baseline \
    > gpurun_out/r2_config3_cap13_exact_g$N.json 2> gpurun_out/r2_config3_cap13_exact_g$N.err
tail -3 gpurun_out/r2_config3_cap13_exact_g$N.err; grep '^{' gpurun_out/r2_config3_cap13_exact_g$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', d['ms_per_step'], 'value', d['value'], 'accuracy', d['accuracy'])
for r,p in enumerate(d['phases_us_all_ranks']): print(r, p)
"
