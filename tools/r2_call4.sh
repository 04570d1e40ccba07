#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rfs -x -k "list or variants or disk_1m" 2>&1 | tail -30 ) > gpurun_out/r2c4_pytest.log
for v in 2 0 2 0; do echo "== bpl $v"; timeout 120 python tools/profile_step.py --warmup 10 --steps 30 --bpl $v 2>&1 | tail -1; done > gpurun_out/r2c4_ab.log 2>&1
for n in 4000000 16000000; do for v in 2 0; do echo "== n $n bpl $v"; timeout 200 python tools/profile_step.py --n $n --warmup 5 --steps 10 --bpl $v 2>&1 | tail -1; done; done >> gpurun_out/r2c4_ab.log 2>&1
for n in 40000 200000; do for v in 1 8; do echo "== n $n bpl $v"; timeout 200 python tools/profile_step.py --n $n --warmup 10 --steps 30 --bpl $v 2>&1 | tail -1; done; done >> gpurun_out/r2c4_ab.log 2>&1
tail -12 gpurun_out/r2c4_pytest.log; cat gpurun_out/r2c4_ab.log
