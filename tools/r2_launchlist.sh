#!/usr/bin/env bash
set -u
for n in 1000000 16000000; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_n$n.csv python tools/profile_step.py --n $n --warmup 2 --steps 2 > gpurun_out/r2_launches_n$n.log 2>&1
done
