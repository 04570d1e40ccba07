#!/usr/bin/env bash
set -u
for n in 1000000 4000000 16000000; do for ps in "" "--presort"; do echo "== n $n $ps"; timeout 300 python tools/profile_step.py --n $n --warmup 10 --steps 20 $ps 2>&1 | tail -1; done; done > gpurun_out/r2_presort.log 2>&1
cat gpurun_out/r2_presort.log
