#!/usr/bin/env bash
# quick GPU check of the traversal kernel: parity subset + warm timing at 1M / 4M (bpl 0 = default list kernel, 2 = pair kernel)
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "list or variants or disk_1m" 2>&1 | tail -8 ) > gpurun_out/quick_pytest.log
for v in 0 2 0; do echo "== bpl $v"; timeout 120 python tools/profile_step.py --warmup 10 --steps 30 --bpl $v 2>&1 | tail -1; done > gpurun_out/quick_ab.log 2>&1
for n in 4000000; do for v in 0; do echo "== n $n bpl $v"; timeout 200 python tools/profile_step.py --n $n --warmup 5 --steps 10 --bpl $v 2>&1 | tail -1; done; done >> gpurun_out/quick_ab.log 2>&1
tail -4 gpurun_out/quick_pytest.log; cat gpurun_out/quick_ab.log | grep -o "==.*\|'traverse_us': [0-9.]*"
