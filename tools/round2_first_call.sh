#!/usr/bin/env bash
# First GPU call of the next round (one B200): everything round 1 left unmeasured, in one pass.
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
# 1. the exact-leaves kernels (BH_FLAG_EXACT_LEAVES) against the oracle extension — never run on a GPU yet;
# 2. the regular GPU suite + smoke;
# 3. bench (default) and the ncu launch list of the SAME command;
# 4. one ncu --set full capture of the production traversal kernel of the final build;
# 5. cost of exact leaves at 1M bodies (profile_step with the flag).
set -u
mkdir -p gpurun_out
( BH_TEST_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_exact_leaves.py tests/test_gpu_generate.py tests/test_gpu_edge_cases.py -m gpu -q -rfs 2>&1 | tail -40 ) > gpurun_out/r2_exact_leaves_pytest.log
( timeout 600 python -m pytest tests -m gpu -q -rfs 2>&1 | tail -30 ) > gpurun_out/r2_pytest.log
( timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/r2_smoke.log
timeout 400 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline \
    > gpurun_out/r2_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_f32_pair -s 2 -c 1 \
    -o gpurun_out/r2_traverse_pair python tools/profile_step.py --steps 4 > gpurun_out/r2_ncu_full.log 2>&1
timeout 120 python tools/profile_step.py --steps 4 > gpurun_out/r2_profile_default.log 2>&1
timeout 120 python tools/profile_step.py --steps 4 --exact-leaves > gpurun_out/r2_profile_exact_leaves.log 2>&1
timeout 120 python tools/profile_step.py --steps 4 --exact-leaves --bpl 2 > gpurun_out/r2_profile_exact_leaves_pair.log 2>&1
# 6. the experimental variants of the pair kernel (bodies_per_lane knob 4 = L1 prefetch, 5 = SM-local block order,
#    6 = both, 7 = pop + fetch of the next cell between test and force phase): parity, then time (traverse_us of each log vs r2_profile_default.log)
( BH_TEST_UNVALIDATED=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k traversal_variants 2>&1 | tail -5 ) > gpurun_out/r2_prefetch_pytest.log
timeout 120 python tools/profile_step.py --steps 6 --bpl 4 > gpurun_out/r2_profile_prefetch.log 2>&1
timeout 120 python tools/profile_step.py --steps 6 --bpl 5 > gpurun_out/r2_profile_sm_local.log 2>&1
timeout 120 python tools/profile_step.py --steps 6 --bpl 6 > gpurun_out/r2_profile_sm_local_prefetch.log 2>&1
timeout 120 python tools/profile_step.py --steps 6 --bpl 7 > gpurun_out/r2_profile_pipelined.log 2>&1
# 8. BASELINE config 1 shape (the reference's own N = 40 000, 10 free-running steps; step 0 is the non-degenerate one)
timeout 120 python tools/free_run.py 40000 10 square > gpurun_out/r2_free_run_40000.log 2>&1
tail -5 gpurun_out/r2_exact_leaves_pytest.log gpurun_out/r2_pytest.log gpurun_out/r2_smoke.log
tail -3 gpurun_out/r2_prefetch_pytest.log; tail -1 gpurun_out/r2_profile_default.log gpurun_out/r2_profile_exact_leaves.log gpurun_out/r2_profile_exact_leaves_pair.log gpurun_out/r2_profile_prefetch.log gpurun_out/r2_profile_sm_local.log gpurun_out/r2_profile_sm_local_prefetch.log gpurun_out/r2_profile_pipelined.log
cut -c1-300 gpurun_out/r2_bench_default.json
# 7. (needs 2 GPUs: gpurun --gpus 2) the pipelined multi-rank host step, parity then e2e:
#   BH_HOST_PIPELINE_MULTI=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#       --master-port 29533 tests/multi_gpu_check.py --bodies 100001 --steps 2 --host-step
#   BH_HOST_PIPELINE_MULTI=1 python -m torch.distributed.run ... bench.py --gpus 2 --steps 100   (compare e2e with the plain run)
