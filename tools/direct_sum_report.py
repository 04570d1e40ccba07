"""BASELINE config 5: direct all-pairs kernel at N = 262 144 (timing, FP32 rate) and the error of the
Barnes-Hut tree forces (theta = 0.5, reference semantics) against it.  The reference's own tree force
includes the self-inclusive cap-leaf term, so errors are reported over all bodies and over
bodies that sit alone in their leaf."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_nbody_simulation_b200 as bh
from gpu_nbody_simulation_b200 import initial_conditions as ic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262_144
pos, vel, mass = ic.uniform_disk(n, seed=12345, round6=False)
with bh.Simulation(n, counters=True) as sim:
    sim.set_bodies(pos, vel, mass)
    sim.direct_forces(want_output=False)                      # warm-up
    fd, ms = sim.direct_forces()
    sim.build_tree(); sim.compute_forces()
    ft = sim.forces()
    c = sim.counters()
    t = sim.tree()
pairs = float(n) * (n - 1)
single = np.zeros(n, dtype=bool)
occ = t[(t[:, 9] == 0) & (t[:, 8] != -1), 8].astype(np.int64)
single[np.where(occ >= 0, occ, -occ - 2)] = True
err = np.linalg.norm(ft - fd, axis=1) / np.maximum(np.linalg.norm(fd, axis=1), 1e-300)
def rms(a, b): return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))
out = {"n": n, "direct_ms": ms, "pair_interactions_per_s": pairs / (ms * 1e-3),
       "direct_tflops_at_20_flop": pairs * 20 / (ms * 1e-3) / 1e12,
       "tree_interactions_per_body": c["interactions"] / n,
       "tree_vs_direct_rel_rms_all": rms(ft, fd), "tree_vs_direct_median_all": float(np.median(err)),
       "single_occupant_bodies": int(single.sum()),
       "tree_vs_direct_rel_rms_single": rms(ft[single], fd[single]),
       "tree_vs_direct_median_single": float(np.median(err[single])),
       "tree_vs_direct_p90_single": float(np.percentile(err[single], 90))}
print(json.dumps(out))
