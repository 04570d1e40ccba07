#!/bin/bash
# The reference's scaling-script shape extended to GPUs (BASELINE config 4): fixed total problem
# size, sweep over the number of B200s of one box, repeats, same results grammar — the "n_threads"
# column holds the GPU count, so plot_first_scale.py's speedup / efficiency plots work unchanged.
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
bodies=(${BODIES:-4000000 16000000 64000000})
gpus=(${GPUS:-1 2 4 8})
steps=${STEPS:-10}
repeats=${REPEATS:-5}
output="gpu_scaling_results.txt"
echo "n_bodies, n_threads, n_simulations, runtime" > $output
for n_b in "${bodies[@]}"; do
    for g in "${gpus[@]}"; do
        for ((i=1; i<=repeats; i++)); do
            if [ "$g" -eq 1 ]; then launcher="python"; else
                launcher="python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g))"; fi
            line=$($launcher "$here/../bench.py" --gpus $g --steps $steps --total-bodies $n_b --quick --reference-lines | tail -1)
            echo "$n_b, $g, $steps, $line" >> $output
            echo "n_bodies=$n_b gpus=$g repeat $i/$repeats: $line"
        done
    done
done
