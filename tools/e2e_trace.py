"""Timeline of bh_step_host (BH_HOST_TRACE=1) for 1 and 4 index chunks, plus raw pinned H2D / D2H rates.
python tools/e2e_trace.py 2> trace.txt"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BH_HOST_TRACE"] = "1"
import gpu_nbody_simulation_b200 as bh  # noqa: E402
from gpu_nbody_simulation_b200 import initial_conditions as ic  # noqa: E402

n = 1_000_000
pos, vel, mass = ic.uniform_disk(n, seed=12345, round6=True)
hp, hv, hm = (torch.from_numpy(x).pin_memory() for x in (pos, vel, mass))
hout = torch.empty((n, 2), dtype=torch.float64).pin_memory()
# raw copy rates, 16 MB and 40 MB, alone and both directions at once
d = torch.empty(40_000_000, dtype=torch.uint8, device="cuda")
h = torch.empty(40_000_000, dtype=torch.uint8).pin_memory()
d2 = torch.empty(16_000_000, dtype=torch.uint8, device="cuda")
h2 = torch.empty(16_000_000, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for label, fn in (("H2D 40MB", lambda: d.copy_(h, non_blocking=True)), ("D2H 16MB", lambda: h2.copy_(d2, non_blocking=True))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(f"[raw] {label}: {dt * 1e3:.3f} ms", file=sys.stderr)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print(f"[raw] H2D 40MB + D2H 16MB concurrently: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms", file=sys.stderr)
for ch in (1, 4):
    os.environ["BH_HOST_CHUNKS"] = str(ch)
    with bh.Simulation(n, device=0) as sim:
        for i in range(3):
            print(f"--- chunks={ch} call {i}", file=sys.stderr)
            sim.step_host(hp, hv, hm, hout)
