"""BASELINE config 3 on ONE GPU: Plummer sphere projected to 2-D, N bodies, raised depth cap, exact in-leaf pairs
(BH_FLAG_EXACT_LEAVES) — throughput, force error against the oracle's exact-leaves tree forces (sampled), and
tree-vs-direct-sum error (sampled), as one JSON line.

    python tools/config3_report.py [--n 16000000] [--max-depth 13] [--steps 5] [--no-exact-leaves]
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_nbody_simulation_b200 as bh  # noqa: E402
import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16_000_000)
ap.add_argument("--max-depth", type=int, default=13)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--no-exact-leaves", action="store_true")
ap.add_argument("--bpl", type=int, default=2, help="2 = pair kernel with the packed member loop (exact leaves)")
ap.add_argument("--samples", type=int, default=1024)
a = ap.parse_args()
exact = not a.no_exact_leaves
pos, vel, mass = bh.generate_host("plummer_2d", a.n, seed=12345)
out = {"config": f"Plummer 2-D N={a.n}, a=0.02, r<=0.1, seed 12345 (bh_generate_host), theta=0.5, depth cap {a.max_depth}, "
                 f"exact_leaves={exact}, one GPU"}
with bh.Simulation(a.n, max_depth=a.max_depth, exact_leaves=exact, bodies_per_lane=a.bpl if exact else 0) as sim:
    sim.set_bodies(pos, vel, mass)
    sim.snapshot()
    sim.step_from_snapshot(2)
    sim.synchronize()
    ms = []
    for _ in range(a.steps):
        sim.step_from_snapshot(1)
        ms.append(sim.last_step_ms())
    out["ms_per_step"] = statistics.median(ms)
    out["body_steps_per_s"] = a.n / (out["ms_per_step"] * 1e-3)
    sim.set_profiling(True)
    sim.step_from_snapshot(1); sim.reset_timers(); sim.step_from_snapshot(2); sim.synchronize()
    t = sim.timers()
    out["phases_us"] = {k: round(t[k] / max(t["steps"], 1), 1) for k in ("bounds_keys_us", "sort_us", "build_us", "traverse_us")}
    sim.set_profiling(False)
    sim.restore()
    sim.build_tree()
    out["tree_nodes"] = sim.tree_size()
    sim.compute_forces()
    f = sim.forces()
t0 = time.perf_counter()
tree = oracle.Tree(pos, mass, oracle.default_params(max_depth=a.max_depth))
out["oracle_tree_build_s"] = round(time.perf_counter() - t0, 1)
assert tree.size == out["tree_nodes"], (tree.size, out["tree_nodes"])
stride = max(1, a.n // a.samples)
fn = tree.forces_exact_leaves if exact else tree.forces
f_ref, cnt = fn(stride=stride, nthreads=oracle.max_threads())
sel = np.arange(0, a.n, stride)
ok = np.isfinite(f_ref[sel]).all(axis=1)
out["interactions_per_body_oracle"] = cnt["interactions"] / len(sel)
out["force_rel_rms_vs_oracle_tree"] = float(np.sqrt(((f[sel][ok] - f_ref[sel][ok]) ** 2).sum() / (f_ref[sel][ok] ** 2).sum()))
dsel = sel[:: max(1, len(sel) // 48)][:48]
fd = np.stack([oracle.direct_forces(pos, mass, i0=int(i), i1=int(i) + 1, nthreads=oracle.max_threads())[int(i)] for i in dsel])
out["tree_vs_direct_sum_rel_rms"] = float(np.sqrt(((f[dsel] - fd) ** 2).sum() / (fd ** 2).sum()))
per = np.linalg.norm(f[dsel] - fd, axis=1) / np.maximum(np.linalg.norm(fd, axis=1), 1e-300)
out["tree_vs_direct_sum_median_per_body"] = float(np.median(per))
out["direct_sum_sample"] = f"{len(dsel)} bodies x all {a.n} partners (FP64, oracle/bh_oracle.c: bho_direct_forces)"
print(json.dumps(out), flush=True)
