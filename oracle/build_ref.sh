#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Builds the reference's own CPU path (project.cu, compiled where it
# lies under /root/reference, unmodified, via oracle/ref_harness.cu) into oracle/_ref/.
# One binary per body count because N_BODIES is a compile-time array size in the reference
# (project.cu:1-3, :38-43).  Outputs go ONLY to oracle/_ref/ (git-ignored, travels with gpurun).
#
#   oracle/build_ref.sh 40000 1000000        -> oracle/_ref/ref_harness_N40000, ..._N1000000
#   oracle/build_ref.sh gpu 1000000 1        -> oracle/_ref/ref_gpu_N1000000_S1 (runSimulationGpu, sm_100a)
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${BH_REFERENCE_ROOT:-/root/reference}/implementation/project.cu"
if [ ! -f "$ref" ]; then
    echo "build_ref: $ref not present (GPU box?) - keeping prebuilt oracle/_ref as is" >&2
    exit 0
fi
mkdir -p "$here/_ref"
if [ "${1:-}" = "direct" ]; then
    # the reference's direct-sum force function (main_approach_1.cpp:53-75; n = 2 is hard-coded there)
    src="${BH_REFERENCE_ROOT:-/root/reference}/implementation/main_approach_1.cpp"
    out="$here/_ref/ref_direct"
    if [ -x "$out" ] && [ "$out" -nt "$here/ref_direct_harness.cpp" ] && [ "$out" -nt "$src" ]; then exit 0; fi
    g++ -O2 -w -std=c++17 -ffp-contract=off -DREF_DIRECT_SOURCE="\"$src\"" -o "$out" "$here/ref_direct_harness.cpp"
    echo "built $out"
    exit 0
fi
if [ "${1:-}" = "approach2" ]; then
    # the reference's stand-alone CPU Barnes-Hut program (main_approach_2.cpp; N_BODIES = 1000 is hard-coded there)
    src="${BH_REFERENCE_ROOT:-/root/reference}/implementation/main_approach_2.cpp"
    out="$here/_ref/ref_approach2"
    if [ -x "$out" ] && [ "$out" -nt "$here/ref_approach2_harness.cpp" ] && [ "$out" -nt "$src" ]; then exit 0; fi
    g++ -O2 -w -std=c++17 -ffp-contract=off -DREF_APPROACH2_SOURCE="\"$src\"" -o "$out" "$here/ref_approach2_harness.cpp"
    echo "built $out"
    exit 0
fi
if [ "${1:-}" = "gpu" ]; then
    # the reference's GPU program path (runSimulationGpu) for the B200 baseline: oracle/build_ref.sh gpu <N> <steps>
    # unmodified project.cu, -O2, sm_100a, N_THREADS = N_BODIES.
    n="$2"; s="$3"
    out="$here/_ref/ref_gpu_N${n}_S${s}"
    if [ -x "$out" ] && [ "$out" -nt "$here/ref_gpu_harness.cu" ] && [ "$out" -nt "$ref" ]; then exit 0; fi
    nvcc -O2 -w -std=c++17 -gencode arch=compute_100a,code=sm_100a -DN_BODIES="$n" -DN_THREADS="$n" \
         -DN_SIMULATIONS="$s" -DREF_SOURCE="\"$ref\"" -o "$out" "$here/ref_gpu_harness.cu" -lpthread
    echo "built $out"
    exit 0
fi
for n in "$@"; do
    out="$here/_ref/ref_harness_N$n"
    if [ -x "$out" ] && [ "$out" -nt "$here/ref_harness.cu" ] && [ "$out" -nt "$ref" ]; then continue; fi
    # -O2 as in SURVEY 8(c); host code only ever executes; static cudart so the binary is self-contained.
    nvcc -O2 -w -std=c++17 -DN_BODIES="$n" -DREF_SOURCE="\"$ref\"" \
         -o "$out" "$here/ref_harness.cu"
    echo "built $out"
done
