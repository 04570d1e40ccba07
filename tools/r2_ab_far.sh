#!/usr/bin/env bash
# A/B of the round-2 closing experiments (one GPU): far class-A nodes without the distance offset (default build),
# the far partial-mask loop unrolled x4 (libbh_u4.so), tree_bottom at 5 blocks per SM (libbh_mb5.so) against the
# validated build (libbh_base.so = -DBH_FAR_NO_EPS=0).  Parity first, then warm phase timers at 1M / 4M.
set -u
mkdir -p gpurun_out
L=$PWD/gpu_nbody_simulation_b200
( for lib in libbh.so libbh_u4.so; do echo "== pytest $lib"
    BH_LIB=$L/$lib timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_reorder.py tests/test_gpu_exact_leaves.py -m gpu -q -x 2>&1 | tail -4
  done ) > gpurun_out/r2f_pytest.log 2>&1
( for rep in 1 2; do for lib in libbh_base.so libbh.so libbh_u4.so libbh_mb5.so; do echo "== 1M $lib"
    BH_LIB=$L/$lib timeout 120 python tools/profile_step.py --warmup 10 --steps 30 2>&1 | tail -1; done; done
  for lib in libbh_base.so libbh.so libbh_u4.so; do echo "== 4M $lib"
    BH_LIB=$L/$lib timeout 200 python tools/profile_step.py --n 4000000 --warmup 5 --steps 10 2>&1 | tail -1; done ) > gpurun_out/r2f_ab.log 2>&1
cat gpurun_out/r2f_pytest.log; grep -o "==.*\|'sort_us.*" gpurun_out/r2f_ab.log
