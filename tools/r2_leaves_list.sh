#!/usr/bin/env bash
# exact leaves in the list kernel (reserved[0] = 9) against the pair kernel's member loop (2): parity, then warm timers
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_exact_leaves.py -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/r2l_pytest.log
( for v in 2 9; do echo "== 1M disk cap 10 bpl $v"; timeout 120 python tools/profile_step.py --exact-leaves --bpl $v --warmup 3 --steps 10 2>&1 | tail -1; done
  for v in 2 9; do echo "== 4M plummer cap 13 bpl $v"; timeout 200 python tools/profile_step.py --n 4000000 --dist plummer --max-depth 13 --exact-leaves --bpl $v --warmup 2 --steps 5 2>&1 | tail -1; done ) > gpurun_out/r2l_ab.log 2>&1
cat gpurun_out/r2l_pytest.log; grep -o "==.*\|'traverse_us': [0-9.]*\|Error.*" gpurun_out/r2l_ab.log
