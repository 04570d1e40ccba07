#!/usr/bin/env bash
# final 2-GPU check of the round: bench.py under torch.distributed.run (weak headline + 16M strong object + accuracy)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --brackets 10 > gpurun_out/r2_bench_g2.json 2> gpurun_out/r2_bench_g2.err
tail -3 gpurun_out/r2_bench_g2.err; grep '^{' gpurun_out/r2_bench_g2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'], 'accuracy', d['accuracy'] and d['accuracy']['force_rel_rms_vs_reference_tree'])
print('roofline', d['roofline'] and d['roofline']['frac'], 'phases', d['phases_us'])
s=d.get('strong'); print('strong', s and {k: s[k] for k in s if k in ('ms_per_step','speedup_vs_1gpu','ms_per_step_1gpu','accuracy')})
"
