"""Free-running steps with the reference's constants (the tree collapses after one step):
robustness / timing of the degenerate regime (huge boxes, almost all bodies in a few finest cells)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_nbody_simulation_b200 as bh
from gpu_nbody_simulation_b200 import initial_conditions as ic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dist = sys.argv[3] if len(sys.argv) > 3 else "disk"
gen = {"disk": ic.uniform_disk, "plummer": ic.plummer_2d, "square": ic.uniform_square}[dist]
pos, vel, mass = gen(n, seed=12345, round6=False)
with bh.Simulation(n, counters=True, graph=False) as sim:
    sim.set_bodies(pos, vel, mass)
    sim.set_profiling(True)
    for s in range(steps):
        t0 = time.perf_counter()
        sim.step(1); sim.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        c = sim.counters()
        t = sim.timers(); sim.reset_timers()
        p = sim.positions()
        print(f"step {s}: {dt:8.3f} ms  nodes {c['nodes']:7d}  heavy cells {c['heavy_cells']:6d}  interactions/body {c['interactions'] / n:8.1f}"
              f"  finite {np.isfinite(p).all()}  |x|max {np.nanmax(np.abs(p)):.3e}"
              f"  us: keys {t['bounds_keys_us']:.0f} sort {t['sort_us']:.0f} build {t['build_us']:.0f} trav {t['traverse_us']:.0f}", flush=True)
