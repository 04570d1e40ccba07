#!/usr/bin/env bash
# One-GPU validation pass (run under gpurun): parity tests, smoke, bench (default paths and the A/B
# fall-backs), the reference GPU program baseline, and the ncu launch list of the bench command.
# Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/smi.txt 2>&1
( timeout 600 python -m pytest tests -m gpu -q -rfs 2>&1 | tail -60; echo "rc=${PIPESTATUS[0]}" ) > gpurun_out/pytest.log
( timeout 120 python __graft_entry__.py smoke 2>&1 | tail -5; echo "smoke rc=${PIPESTATUS[0]}" ) > gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
BH_PDL=1 timeout 200 python bench.py --steps 100 --quick > gpurun_out/bench_pdl1.log 2>&1
BH_REORDER=0 timeout 200 python bench.py --steps 100 --quick > gpurun_out/bench_reorder0.log 2>&1
timeout 200 python tools/e2e_sweep.py > gpurun_out/e2e_sweep.log 2>&1
if [ "${1:-}" = "ncu" ]; then
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
        --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --brackets 1 --quick \
        > gpurun_out/ncu_bench.log 2>&1
fi
tail -25 gpurun_out/pytest.log; cat gpurun_out/smoke.log; cat gpurun_out/e2e_sweep.log; rm -f gpurun_out/bench_oldpaths.log gpurun_out/bench_snapcopy.log gpurun_out/bench_bisect.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.log")):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            r = d.get("roofline") or {}
            print(f.split("/")[-1], round(d["value"] / 1e9, 3), "G/s", round(d["ms_per_step"], 4), "ms e2e",
                  round(d["e2e"]["value"] / 1e9, 3), d.get("phases_us"), (d.get("gpu_baseline") or {}).get("value"),
                  d["clocks"])
PY
