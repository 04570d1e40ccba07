#!/usr/bin/env bash
# ncu --set full of the seven build-phase kernels of one warm step (bounds, keys, 2 sort passes, cell scan, heavy cells,
# tree bottom) at N = 1M and N = 16M (uniform disk, resident order): the HBM / issue view of everything but the traversal.
set -u
mkdir -p gpurun_out
for n in 1000000 16000000; do
    timeout 500 ncu --set full --clock-control none -k regex:'bounds_kernel|keys_kernel|onesweep_pass|cell_scan|heavy_huge|tree_bottom' \
        -s 21 -c 7 -f -o gpurun_out/r2b_build_$n python tools/profile_step.py --n $n --steps 6 --presort > gpurun_out/r2b_build_$n.log 2>&1
    tail -2 gpurun_out/r2b_build_$n.log
done
