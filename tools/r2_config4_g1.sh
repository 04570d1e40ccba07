#!/usr/bin/env bash
# BASELINE config 4 on ONE GPU with the final build: N = 4M / 16M / 64M uniform disk (bench.py --quick --no-e2e)
set -u
mkdir -p gpurun_out
for N in 4000000 16000000 64000000; do
    steps=$(( N >= 64000000 ? 10 : 30 ))
    timeout 400 python bench.py --gpus 1 --steps $steps --warmup 3 --total-bodies $N --quick --no-e2e \
        > "gpurun_out/r2_strong_${N}_g1.json" 2> "gpurun_out/r2_strong_${N}_g1.err"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_strong_*_g1.json")):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f.split("/")[-1], d["n_gpus"], "GPUs", round(d["value"] / 1e9, 3), "G body-steps/s", round(d["ms_per_step"], 4), "ms/step",
                  {k: round(v, 1) for k, v in d["phases_us"].items()}, "roofline", round(d["roofline"]["frac"], 4), "accuracy",
                  d["accuracy"].get("force_rel_rms_vs_reference_tree"), d["clocks"])
PY
