#!/usr/bin/env bash
# one ncu --set full capture of the production traversal kernel at N = 1M: tools/r2_ncu.sh <tag> [kernel regex]
set -u
tag=${1:-x}; k=${2:-list_kernel}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/r2_$tag python tools/profile_step.py --steps 5 > gpurun_out/r2_ncu_$tag.log 2>&1
tail -1 gpurun_out/r2_ncu_$tag.log
