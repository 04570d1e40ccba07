#!/usr/bin/env bash
# Warm A/B of the traversal variants (bh_params.reserved[0]) at N = 1M: 10 untimed + 30 profiled steps each.
set -u
mkdir -p gpurun_out
for v in 0 4 5 6 7 0 5; do
  echo "== bpl $v"; timeout 120 python tools/profile_step.py --warmup 10 --steps 30 --bpl $v 2>&1 | tail -1
done > gpurun_out/r2_ab_traverse.log 2>&1
for n in 4000000 16000000; do for v in 0 5; do
  echo "== n $n bpl $v"; timeout 200 python tools/profile_step.py --n $n --warmup 5 --steps 10 --bpl $v 2>&1 | tail -1
done; done >> gpurun_out/r2_ab_traverse.log 2>&1
cat gpurun_out/r2_ab_traverse.log
