#!/usr/bin/env bash
# GPU suite (fast subset) + warm phase timing at 1M / 16M
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/quick2_pytest.log
for n in 1000000 16000000; do echo "== n $n"; timeout 300 python tools/profile_step.py --n $n --warmup 10 --steps 20 2>&1 | tail -1; done > gpurun_out/quick2_ab.log 2>&1
tail -4 gpurun_out/quick2_pytest.log; cat gpurun_out/quick2_ab.log
