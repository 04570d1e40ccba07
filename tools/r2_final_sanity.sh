#!/usr/bin/env bash
# last GPU pass of the round: the driver's own sequence (pytest -x, smoke) + the default exact-leaves path at 1M vs the oracle
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2z_pytest.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r2z_smoke.log
timeout 300 python bench.py --exact-leaves --quick --no-e2e --steps 5 --warmup 3 --brackets 5 > gpurun_out/r2z_bench_exact_1M.json 2> gpurun_out/r2z_bench_exact_1M.err
cat gpurun_out/r2z_pytest.log gpurun_out/r2z_smoke.log; tail -2 gpurun_out/r2z_bench_exact_1M.err
grep '^{' gpurun_out/r2z_bench_exact_1M.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('exact leaves 1M: ms/step', d['ms_per_step'], 'phases', d['phases_us'], 'accuracy', d['accuracy'])"
