"""Default exact-leaves path (list kernel from 500k bodies on) at N = 1M uniform disk: sampled force error against the
oracle's exact-leaves tree forces (bench.sampled_accuracy), and against the pair kernel's member loop (all bodies).
    python tools/exact_leaves_accuracy.py [n]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import gpu_nbody_simulation_b200 as bh  # noqa: E402
import oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pos, vel, mass = bench.make_workload(n)
out = {"n_bodies": n}
forces = {}
for bpl in (0, 2):
    with bh.Simulation(n, exact_leaves=True, bodies_per_lane=bpl) as sim:
        sim.set_bodies(pos, vel, mass)
        sim.build_tree()
        sim.compute_forces()
        forces[bpl] = sim.forces()
out["default_vs_oracle"] = bench.sampled_accuracy(bh, oracle, forces[0], pos, mass, 10, 0, n, exact_leaves=True)
ok = np.isfinite(forces[2]).all(axis=1)
out["default_vs_pair_kernel_rel_rms_all_bodies"] = float(np.sqrt(np.sum((forces[0][ok] - forces[2][ok]) ** 2) / np.sum(forces[2][ok] ** 2)))
print(json.dumps(out))
