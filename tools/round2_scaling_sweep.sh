#!/usr/bin/env bash
# BASELINE config 4 (strong scaling N = 4M / 16M / 64M) and config 3 (Plummer 16M, cap 10 and a raised cap) for ONE
# GPU count, so that each count can run in its own gpurun call (charged N x the box time):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_scaling_sweep.sh 8'
#   gpurun --gpus 1 --timeout 900 -- 'bash tools/round2_scaling_sweep.sh 1'      (likewise 2, 4)
# Results: gpurun_out/r2_strong_<N>_g<G>.json, r2_plummer16M_cap<D>_g<G>.json
set -u
G="${1:-1}"
mkdir -p gpurun_out
if [ "$G" -eq 1 ]; then L="python"; else
    L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29600 + G))"; fi
for N in 4000000 16000000 64000000; do
    steps=$(( N >= 64000000 ? 10 : 30 ))
    timeout 400 $L bench.py --gpus "$G" --steps $steps --warmup 3 --total-bodies $N --quick --no-e2e \
        > "gpurun_out/r2_strong_${N}_g${G}.json" 2> "gpurun_out/r2_strong_${N}_g${G}.err"
done
for D in 10 12; do
    timeout 400 $L bench.py --gpus "$G" --steps 20 --warmup 3 --total-bodies 16000000 --dist plummer --max-depth $D \
        --quick --no-e2e > "gpurun_out/r2_plummer16M_cap${D}_g${G}.json" 2> "gpurun_out/r2_plummer16M_cap${D}_g${G}.err"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_strong_*_g*.json") + glob.glob("gpurun_out/r2_plummer16M_*_g*.json")):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f.split("/")[-1], d["n_gpus"], "GPUs", round(d["value"] / 1e9, 3), "G body-steps/s", round(d["ms_per_step"], 4), "ms/step",
                  d["phases_us"], d["clocks"])
PY
