"""The reference's own GPU program (unmodified project.cu, oracle/_ref/ref_gpu_N*_S*) on this GPU:
its two timers for the shipped 40 000 bodies and the 1M disk, 1 step and the reference's 10 steps.
Prints one JSON object.  Bench / test infrastructure (uses oracle/)."""
import json
import os
import statistics
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from gpu_nbody_simulation_b200 import initial_conditions as ic  # noqa: E402

out = {}
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "shipped_40000.npz"))
sets = {40000: (g["pos"], g["vel"], g["mass"]), 1000000: ic.uniform_disk(1_000_000, seed=12345, round6=True)}
for n, (pos, vel, mass) in sets.items():
    for steps in (1, 10):
        key = f"N{n}_S{steps}"
        if not oracle.ref_gpu_available(n, steps):
            out[key] = {"unavailable": "binary not built"}
            continue
        try:
            calls, _ = oracle.run_ref_gpu(pos, vel, mass, steps=steps, calls=4, timeout=150)
        except Exception as e:
            out[key] = {"unavailable": str(e)[:300]}
            continue
        timed = calls[1:]
        tot = statistics.median(c["total_ms"] for c in timed)
        par = statistics.median(c["parallel_us"] for c in timed)
        out[key] = {"total_ms": tot, "parallel_us": par, "body_steps_per_s_total": n * steps / (tot * 1e-3),
                    "body_steps_per_s_kernels": n * steps / (par * 1e-6), "calls": calls}
print(json.dumps(out))
