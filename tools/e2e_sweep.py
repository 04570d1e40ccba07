"""A/B of the pipelined host step (bh_step_host) on one GPU: number of index chunks (BH_HOST_CHUNKS).
Prints one JSON line per setting.  python tools/e2e_sweep.py [--n 1000000] [--steps 40]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_nbody_simulation_b200 as bh  # noqa: E402
from gpu_nbody_simulation_b200 import initial_conditions as ic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--chunks", default="1,2,3,4,6")
a = ap.parse_args()
pos, vel, mass = ic.uniform_disk(a.n, seed=12345, round6=a.n <= 2_000_000)
hp, hv, hm = (torch.from_numpy(x).pin_memory() for x in (pos, vel, mass))
hout = torch.empty((a.n, 2), dtype=torch.float64).pin_memory()
ref = None
for ch, graph in [(int(x), g) for g in (True, False) for x in a.chunks.split(",")]:
    os.environ["BH_HOST_CHUNKS"] = str(ch)
    with bh.Simulation(a.n, device=0, graph=graph) as sim:
        for _ in range(3):
            sim.step_host(hp, hv, hm, hout)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            sim.step_host(hp, hv, hm, hout)
        dt = (time.perf_counter() - t0) / a.steps
        out = hout.numpy().copy()
        if ref is None:
            ref = out
        print(json.dumps({"host_chunks": ch, "graph": graph, "ms_per_step": dt * 1e3, "body_steps_per_s": a.n / dt,
                          "device_ms": sim.last_step_ms(), "bit_identical_to_first": bool(np.array_equal(out, ref, equal_nan=True))}),
              flush=True)
