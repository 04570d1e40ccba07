#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -rfs -x 2>&1 | tail -30 ) > gpurun_out/r2c3_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c3_bench_ref.json 2> gpurun_out/r2c3_bench_ref.err
tail -15 gpurun_out/r2c3_pytest.log; tail -3 gpurun_out/r2c3_bench.err; cut -c1-400 gpurun_out/r2c3_bench.json; cut -c1-600 gpurun_out/r2c3_bench_ref.json
