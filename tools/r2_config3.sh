#!/usr/bin/env bash
set -u
timeout 900 python tools/config3_report.py --n 16000000 --max-depth 13 --steps 5 > gpurun_out/r2_config3_cap13_exact.json 2> gpurun_out/r2_config3_cap13_exact.err
timeout 600 python tools/config3_report.py --n 16000000 --max-depth 10 --steps 5 --no-exact-leaves --samples 512 > gpurun_out/r2_config3_cap10_ref.json 2> gpurun_out/r2_config3_cap10_ref.err
tail -3 gpurun_out/r2_config3_cap13_exact.err; cat gpurun_out/r2_config3_cap13_exact.json; tail -3 gpurun_out/r2_config3_cap10_ref.err; cat gpurun_out/r2_config3_cap10_ref.json
