"""Small driver for ncu / compute-sanitizer: a few whole steps of the hot path, direct launches.

    python tools/profile_step.py [--n 1000000] [--steps 3] [--dist disk|plummer|square] [--fp64] [--counters]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_nbody_simulation_b200 as bh  # noqa: E402
from gpu_nbody_simulation_b200 import initial_conditions as ic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--dist", default="disk")
ap.add_argument("--fp64", action="store_true")
ap.add_argument("--counters", action="store_true")
ap.add_argument("--max-depth", type=int, default=10)
ap.add_argument("--bpl", type=int, default=0)
ap.add_argument("--exact-eps", action="store_true")
ap.add_argument("--exact-leaves", action="store_true", help="BH_FLAG_EXACT_LEAVES (extension)")
ap.add_argument("--presort", action="store_true", help="hand the bodies over in Morton order of the initial distribution")
ap.add_argument("--warmup", type=int, default=0, help="untimed steps before the profiled ones (A/B timing: use >= 10)")
a = ap.parse_args()
gen = {"disk": ic.uniform_disk, "plummer": ic.plummer_2d, "square": ic.uniform_square}[a.dist]
pos, vel, mass = gen(a.n, seed=12345, round6=False)
if a.presort:
    import numpy as np
    with bh.Simulation(a.n, max_depth=a.max_depth) as tmp:
        tmp.set_bodies(pos, vel, mass)
        tmp.build_tree()
        order = tmp.sorted_order().astype(np.int64)
    pos, vel, mass = np.ascontiguousarray(pos[order]), np.ascontiguousarray(vel[order]), np.ascontiguousarray(mass[order])
with bh.Simulation(a.n, graph=False, fp64=a.fp64, counters=a.counters, max_depth=a.max_depth, bodies_per_lane=a.bpl, exact_eps=a.exact_eps,
                   exact_leaves=a.exact_leaves) as sim:
    sim.set_bodies(pos, vel, mass)
    sim.snapshot()
    sim.set_profiling(True)
    if a.warmup:
        sim.step_from_snapshot(a.warmup)
        sim.synchronize()
        sim.reset_timers()
    sim.step_from_snapshot(a.steps)
    sim.synchronize()
    t = sim.timers()
    print({k: round(v / max(t["steps"], 1), 1) if k.endswith("_us") else v for k, v in t.items()})
    if a.counters:
        print(sim.counters())
