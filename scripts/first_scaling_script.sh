#!/bin/bash
# Drop-in equivalent of the reference's implementation/first_scaling_script.sh: fixed problem size,
# N_THREADS sweep, 5 repeats, 10 steps; same results-file grammar, so the reference's
# plot_first_scale.py reads first_scaling_results.txt unchanged.  N_THREADS is accepted by the
# B200 engine's project.cu but carries no meaning there: the sweep is kept so that the
# reference workflow runs as is.  Run in a directory holding the three *_init.txt files.
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
src="${BH_PROJECT_CU:-$here/../gpu_nbody_simulation_b200/cli/project.cu}"
bodies=(${BODIES:-40000})
threads=(${THREADS:-1 2 4 8 16 32 64 128 256 512 1024 2048 4096 8192 16384 32768 40000})
simulations=(${SIMULATIONS:-10})
repeats=${REPEATS:-5}
output="first_scaling_results.txt"
echo "n_bodies, n_threads, n_simulations, runtime" > $output
for n_b in "${bodies[@]}"; do
    for n_t in "${threads[@]}"; do
        for ((i=1; i<=repeats; i++)); do
            for n_s in "${simulations[@]}"; do
                echo -e "\n\n=============================================="
                echo "Running simulation with:"
                echo "  n_bodies=$n_b"
                echo "  n_threads=$n_t (repeat $i/$repeats)"
                echo "  n_simulations=$n_s"
                echo "==============================================\n"
                # the reference's compile line (first_scaling_script.sh:30) plus the sm_100a target
                nvcc -gencode arch=compute_100a,code=sm_100a -O3 -diag-suppress 550 \
                     -DN_BODIES=$n_b -DN_THREADS=$n_t -DN_SIMULATIONS=$n_s -o project "$src"
                runtime=$(./project)
                echo "$n_b, $n_t, $n_s, $runtime" >> $output
                echo -e "\n----------------------------------------------"
                echo "Completed: n_bodies=$n_b, n_threads=$n_t, n_simulations=$n_s"
                echo "  Runtime: $runtime ms"
                echo "----------------------------------------------\n"
            done
        done
    done
done
