#!/usr/bin/env bash
# round-2 evidence pass on one B200: GPU suite, smoke, bench, its ncu launch list, one ncu --set full of the traversal kernel
# (and of one sort pass with "sort" as the first argument)
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -rfs 2>&1 | tail -8 ) > gpurun_out/r2p_pytest.log
( timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3; echo "smoke rc=${PIPESTATUS[0]}" ) > gpurun_out/r2p_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2p_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --brackets 1 --quick > gpurun_out/r2p_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:list_kernel -s 6 -c 1 -f -o gpurun_out/r2p_list_final \
    python tools/profile_step.py --steps 8 --presort > gpurun_out/r2p_ncu_full.log 2>&1
if [ "${1:-}" = "sort" ]; then
    timeout 400 ncu --set full --clock-control none -k regex:onesweep_pass -s 6 -c 1 -f -o gpurun_out/r2p_sort_final \
        python tools/profile_step.py --steps 8 --presort > gpurun_out/r2p_ncu_sort.log 2>&1
fi
tail -5 gpurun_out/r2p_pytest.log; cat gpurun_out/r2p_smoke.log; cut -c1-200 gpurun_out/r2p_bench.json; tail -2 gpurun_out/r2p_ncu_full.log
