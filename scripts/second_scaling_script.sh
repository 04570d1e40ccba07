#!/bin/bash
# Drop-in equivalent of the reference's implementation/second_scaling_script.sh ("weak" sweep:
# n_bodies == n_threads); same results-file grammar for plot_second_scale.py.
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
src="${BH_PROJECT_CU:-$here/../gpu_nbody_simulation_b200/cli/project.cu}"
bodies=(${BODIES:-2 4 8 16 32 64 128 256 512 1024 2048 4096 8192 16384 32768 40000})
threads=("${bodies[@]}")
simulations=(${SIMULATIONS:-10})
repetitions=${REPEATS:-5}
output="second_scaling_results.txt"
echo "n_bodies, n_threads, n_simulations, repetition, runtime" > $output
for i in "${!bodies[@]}"; do
    n_b=${bodies[$i]}
    n_t=${threads[$i]}
    n_s=${simulations[0]}
    for ((rep=1; rep<=repetitions; rep++)); do
        echo -e "\n\n=============================================="
        echo "Running simulation with:"
        echo "  n_bodies=$n_b"
        echo "  n_threads=$n_t"
        echo "  n_simulations=$n_s"
        echo "  repetition=$rep"
        echo "==============================================\n"
        nvcc -gencode arch=compute_100a,code=sm_100a -O3 -diag-suppress 550 \
             -DN_BODIES=$n_b -DN_THREADS=$n_t -DN_SIMULATIONS=$n_s -o project "$src"
        runtime=$(./project)
        echo "$n_b, $n_t, $n_s, $rep, $runtime" >> $output
        echo -e "\n----------------------------------------------"
        echo "Completed: n_bodies=$n_b, n_threads=$n_t, n_simulations=$n_s, repetition=$rep"
        echo "  Runtime: $runtime ms"
        echo "----------------------------------------------\n"
    done
done
