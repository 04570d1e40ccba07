#!/usr/bin/env python
"""Benchmark of the Barnes-Hut hot path (BASELINE.json: body·steps/s, theta = 0.5, 2-D).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] — 1 000 000 bodies, uniform disk R = 0.1, seed
12345, the reference's constants (G, dt = 1, theta = 0.5, depth cap 10), drawn by the library's own
counter-based generator (bh_generate_host: the same bodies for every arm and every N; no text round
trip).  A *step* is one pass of the whole hot path: bounds -> cell keys -> radix sort -> tree + COM ->
traversal -> integrate.  Because the reference's own physics flings bodies away after ONE step and
the tree collapses to a few hundred nodes, every step starts from the initial
distribution (a device snapshot that the step reads out of place; nothing is skipped: bounds, keys,
sort, tree, traversal and integrator all run every step) — i.e. every timed step does the full-size,
non-degenerate work of the reference's step 0.  At N > 1 GPUs the body count grows with N (weak
scaling, 1M bodies per GPU, contiguous Morton slices, sharded tree build; the ranks exchange the
bounding box and the per-cell sums over NVLink peer memory, not the bodies).

`value`  : device-resident throughput, inputs in HBM: W warm-up steps, then brackets of EXACTLY K steps
           between two barrier + torch.cuda.synchronize() fences, device time from CUDA events on the
           library's stream, max over ranks; the bracket is repeated (`brackets`) and the MEDIAN is
           reported, so a short K is not an 11 ms sample.  The per-rank working set (~150 MB) exceeds
           the 126 MB L2; `value_l2_flushed` repeats K steps with an explicit L2 flush before each.
`e2e`    : same step through the C-ABI with HOST buffers: pinned H2D of positions, velocities and
           masses + step + D2H of positions every step, wall clock around the synchronous calls.
`roofline`: traversal kernel, 20 flop per accepted interaction against the FP32
           FMA peak measured by a register-resident FMA loop on the same device.
`strong` : (N > 1) BASELINE config 4 / north_star's target: 16M bodies TOTAL on the N GPUs, against the
           same 16M bodies on ONE GPU (rank 0) measured in the same job.
`direct` : (N = 1) BASELINE config 5: the tiled all-pairs kernel at 262 144 bodies.
`cpu_baseline`: the reference's OWN CPU functions (oracle/_ref, 1 thread: it has no threading) on the same
           bodies (+ `port_all_cores`: the oracle port with OpenMP over bodies, labelled as a port).
`gpu_baseline`: the reference's OWN GPU program (unmodified project.cu, sm_100a) on the same GPU, its two timers.
`accuracy`: force rel-RMS of the timed configuration against the reference tree forces (sampled bodies).
`--impl reference`: the CPU reference arm alone, same JSON line shape, same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BODIES_PER_GPU = 1_000_000
STRONG_TOTAL = 16_000_000
DIRECT_N = 262_144
SEED = 12345
FLOP_PER_INTERACTION = 20.0
METRIC = "body_steps_per_s"
UNIT = "body·steps/s"
GEN_KIND = {"disk": "uniform_disk", "plummer": "plummer_2d", "square": "uniform_square"}
# dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum of the production traversal kernel, one
# `ncu --set full` capture at N = 1M uniform disk; see profiles/ (file named in NCU_SOURCE)
NCU_SOURCE = "profiles/r02_traverse_list_ncu_summary.txt"
TRAVERSE_DRAM_BYTES_NCU = 61_915_136 + 25_891_840
TRAVERSE_WARP_INSTRUCTIONS_NCU = 218_505_556


def make_workload(n, dist="disk"):
    """The timed bodies: the library's seeded counter-based generator on the host (csrc/generate.cu, needs no
    GPU), raw FP64 draws — ONE generator and ONE family of workloads for every N, both arms and the CLI."""
    import gpu_nbody_simulation_b200 as bh
    return bh.generate_host(GEN_KIND[dist], n, seed=SEED)


def workload_string(n, world, dist="disk", max_depth=10):
    name = {"disk": "uniform disk", "plummer": "Plummer sphere projected to 2-D (a=0.02, r<=0.1)", "square": "uniform square"}[dist]
    return (f"{name} N={n} ({n // world} per GPU), R=0.1, seed {SEED} (bh_generate_host), theta=0.5, G=6.67e-11, dt=1, "
            f"depth cap {max_depth}; every step restarts from the initial distribution (the full-size work of the "
            "reference's step 0)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self, wait_s=20.0):
        """Starts ONE long-running nvidia-smi (loop mode) and returns only after its first sample has
        arrived: nvidia-smi's start-up attaches to every GPU of the box and takes 1-2 s on an 8-GPU
        node, during which kernel launches of all ranks stall for milliseconds — started right before
        the timed bracket (as an earlier version did, once per rank) it cost the 4- and 8-GPU runs
        30-60 %.  Polling afterwards is cheap."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < wait_s and self.proc.poll() is None:
                time.sleep(0.02)
        except OSError:
            self.proc = None

    def mark(self):
        """Index of the next sample: samples from here on were taken inside the timed region."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines[getattr(self, "first", 0):]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def numa_node_of_gpu(dev):
    """(node, reason): the NUMA node the GPU hangs off according to sysfs, or (None, why not)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(dev)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception as e:
        return None, f"no PCI address for device {dev}: {type(e).__name__}"
    path = f"/sys/bus/pci/devices/{bdf}/numa_node"
    try:
        node = int(open(path).read())
    except OSError:
        return None, f"{path} not readable (container without the PCI sysfs tree)"
    if node < 0:
        return None, f"{path} = {node}: the platform exposes no NUMA affinity for this GPU (single node or virtualised)"
    return node, "sysfs"


def pin_to_gpu_numa_node(dev):
    """Binds this process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that the pinned
    host buffers of the e2e leg are allocated next to the GPU's PCIe root.  Best effort: returns the
    node number or None (no sysfs entry, single node, restricted cpuset)."""
    return pin_to_gpu_numa_node_why(dev)[0]


def probe_best_numa_node(dev, mb=64):
    """When sysfs does not name the GPU's node: measure it.  For every NUMA node with CPUs in this process's cpuset, bind
    to its CPUs, allocate a pinned buffer there (first touch) and time a host-to-device copy; the fastest node wins.
    Returns (node, {node: GB/s}) or (None, reason)."""
    import glob
    import torch
    nodes = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        try:
            cpus = set()
            for part in open(os.path.join(d, "cpulist")).read().strip().split(","):
                if part:
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                nodes[int(os.path.basename(d)[4:])] = cpus
        except (OSError, ValueError):
            continue
    if len(nodes) < 2:
        return None, f"{len(nodes)} NUMA node(s) with CPUs of this cpuset: nothing to choose"
    before = os.sched_getaffinity(0)
    rates = {}
    dst = torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{dev}")
    for node, cpus in nodes.items():
        os.sched_setaffinity(0, cpus)
        src = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
        src.fill_(1)                                             # first touch on this node
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        rates[node] = 4 * (mb << 20) / (time.perf_counter() - t0) / 1e9
        del src
    os.sched_setaffinity(0, before)
    best = max(rates, key=rates.get)
    return best, {k: round(v, 1) for k, v in rates.items()}


def pin_to_gpu_numa_node_why(dev):
    node, why = numa_node_of_gpu(dev)
    if node is None:
        try:
            node, rates = probe_best_numa_node(dev)
        except Exception as e:
            node, rates = None, f"probe failed: {type(e).__name__}: {e}"
        if node is None:
            return None, f"{why}; probe: {rates}"
        why = f"sysfs names no node; measured H2D GB/s per node {rates}"
    try:
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None, f"node {node}: none of its CPUs is in this process's cpuset"
        os.sched_setaffinity(0, cpus)
        return node, f"bound to the {len(cpus)} CPUs of node {node} ({why})"
    except Exception as e:
        return None, f"node {node}: {type(e).__name__}: {e}"


def cpu_reference_run(n_bodies, steps, warmup, budget_s=150.0):
    """Times the reference's OWN CPU path (oracle/_ref, unmodified project.cu functions, 1 thread — the
    reference has no threading) on the first `n_bodies` bodies of the bench workload (the whole 1M workload
    when n_bodies = 1M), every step restarted from the initial distribution like the GPU arm.  The body count is
    NEVER reduced silently: if the binary for n_bodies is missing the oracle port is timed instead (kind "port").
    A step of the 1M workload takes ~4 s, so when `steps + warmup` full steps do not fit into `budget_s` fewer
    steps are executed (each still the full workload; the rate is per step) and `steps_executed` says so.
    Returns (value, info)."""
    import oracle
    pos, vel, mass = make_workload(BODIES_PER_GPU)
    pos, vel, mass = pos[:n_bodies], vel[:n_bodies], mass[:n_bodies]
    if oracle.ref_available(n_bodies):
        t0 = time.perf_counter()
        _, first = oracle.run_ref(pos, vel, mass, steps=1, dump="", keep_dump=False, reset_each_step=True)
        probe_wall = time.perf_counter() - t0                       # includes writing / reading the 40 MB input
        per_step = sum(first[0][k] for k in ("build_us", "force_us", "update_us")) * 1e-6
        afford = int(max(0.0, budget_s - probe_wall) / max(per_step, 1e-9))
        want = steps + max(0, warmup - 1)                            # the probe step was the first warm-up step
        run = max(1, min(want, afford))
        _, tim = oracle.run_ref(pos, vel, mass, steps=run, dump="", keep_dump=False, reset_each_step=True)
        timed = tim[min(max(0, warmup - 1), run - 1):] if run > steps else tim
        timed = timed[-steps:]
        secs = sum(t["build_us"] + t["force_us"] + t["update_us"] for t in timed) * 1e-6
        value = n_bodies * len(timed) / secs
        info = {"value": value, "unit": UNIT, "cores": 1, "kind": "reference", "n_bodies": n_bodies,
                "steps_executed": len(timed),
                "sample": (f"{len(timed)} full steps (buildTree + computeForces + update*) of the reference's own CPU "
                           f"functions (oracle/_ref/ref_harness_N{n_bodies}: unmodified project.cu, -O2, 1 thread) on "
                           f"{'all' if n_bodies == BODIES_PER_GPU else 'the first'} {n_bodies} bodies of the bench workload, "
                           f"restarted from the initial distribution every step; build "
                           f"{sum(t['build_us'] for t in timed) / len(timed) / 1e3:.0f} ms, force "
                           f"{sum(t['force_us'] for t in timed) / len(timed) / 1e3:.0f} ms per step"
                           + ("" if len(timed) == steps else f"; {steps} steps asked, {len(timed)} fit the {budget_s:.0f} s budget")),
                "ms_per_step": secs / len(timed) * 1e3}
        return value, info
    # oracle port: full tree build, forces on a strided subset, single thread
    stride = 64
    t0 = time.perf_counter()
    tree = oracle.Tree(pos, mass)
    t1 = time.perf_counter()
    tree.forces(stride=stride, nthreads=1)
    t2 = time.perf_counter()
    secs = (t1 - t0) + (t2 - t1) * stride
    value = n_bodies / secs
    return value, {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "n_bodies": n_bodies, "steps_executed": 1,
                   "sample": f"oracle port (oracle/_ref/ref_harness_N{n_bodies} not built): full {n_bodies}-body tree build + "
                             f"forces for every {stride}th body, extrapolated",
                   "ms_per_step": secs * 1e3}


def cpu_port_all_cores(pos, mass, budget_s=8.0):
    """The oracle PORT (oracle/bh_oracle.c, OpenMP over bodies in the force loop — the modification SURVEY 8d allows
    as a clearly labelled extra; the reference itself has no threading) on all host cores: serial tree build +
    forces for every `stride`-th body, extrapolated to all bodies.  Bounded to ~budget_s seconds."""
    import oracle
    n = mass.shape[0]
    nt = oracle.max_threads()
    t0 = time.perf_counter()
    tree = oracle.Tree(pos, mass)
    t1 = time.perf_counter()
    probe_stride = 256
    tree.forces(stride=probe_stride, nthreads=nt)                 # short probe to size the sample
    est_full = (time.perf_counter() - t1) * probe_stride
    stride = max(1, int(est_full / budget_s + 0.999))
    t2 = time.perf_counter()
    tree.forces(stride=stride, nthreads=nt)
    t3 = time.perf_counter()
    secs = (t1 - t0) + (t3 - t2) * stride
    return {"value": n / secs, "unit": UNIT, "cores": nt, "kind": "port (OpenMP over bodies, not the reference's code path)",
            "sample": f"serial tree build {1e3 * (t1 - t0):.0f} ms + forces for every {stride}th of {n} bodies on {nt} threads "
                      f"({1e3 * (t3 - t2):.0f} ms), extrapolated; integrator not included (bandwidth-trivial)"}


def reference_gpu_run(pos, vel, mass, device):
    """The reference's OWN GPU program path on this GPU (north_star's second baseline): unmodified
    project.cu, runSimulationGpu with N_THREADS = N_BODIES, compiled for sm_100a (oracle/_ref/
    ref_gpu_N1000000_S1: one step per call = the non-degenerate step 0, host tree build + tree H2D +
    force and update kernels + positions D2H), timed with the reference's own two timers.  Call 0 pays
    the CUDA context creation and is dropped."""
    import oracle
    n = mass.shape[0]
    if not oracle.ref_gpu_available(n, 1):
        return {"unavailable": f"oracle/_ref/ref_gpu_N{n}_S1 not built"}
    try:
        calls, _ = oracle.run_ref_gpu(pos, vel, mass, steps=1, calls=4, device=device, timeout=300)
    except Exception as e:   # the baseline must never take the bench line down with it
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    timed = calls[1:]
    total_ms = statistics.median(c["total_ms"] for c in timed)
    par_us = statistics.median(c["parallel_us"] for c in timed)
    return {"value": n / (total_ms * 1e-3), "unit": UNIT, "kind": "reference project.cu (unmodified, runSimulationGpu, "
            "N_THREADS = N_BODIES, nvcc -O2 sm_100a) on this GPU",
            "total_ms_per_step": total_ms, "gpu_parallel_us_per_step": par_us,
            "value_kernels_only": n / (par_us * 1e-6),
            "sample": f"median of {len(timed)} calls of runSimulationGpu with N_SIMULATIONS = 1 on the same {n} bodies (step 0, "
                      f"{timed[-1]['last_tree_nodes']} tree nodes); total = the reference's 'GPU total computation' timer "
                      "(mallocs, host tree build, quadtree_init_gpu.txt, copies), gpu_parallel = its 'GPU parallel "
                      "computation' timer (force + update kernels)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = BODIES_PER_GPU
    value, info = cpu_reference_run(n, args.steps, args.warmup)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "steps_executed": info["steps_executed"],
            "warmup": args.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_string(n, 1), "n_bodies": info["n_bodies"],
                       "note": "CPU reference arm: always the N = 1M single-GPU workload, whatever --gpus says (a step of "
                               "the N-GPU weak workload would take N x 4 s); see cpu_baseline.sample"},
            "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bracket_plan(K, target_steps=1000, lo=5, hi=50):
    """Number of K-step brackets: enough for ~target_steps timed steps in total, between lo and hi."""
    return int(max(lo, min(hi, -(-target_steps // max(K, 1)))))


def sampled_accuracy(bh, oracle, f_gpu, pos, mass, max_depth, lo, hi, nsample=4096, exact_leaves=False):
    """Force rel-RMS of f_gpu[lo:hi] against the reference tree forces (pinned oracle) on ~nsample bodies of [lo, hi)."""
    stride = max(1, (hi - lo) // nsample)
    tree = oracle.Tree(pos, mass, oracle.default_params(max_depth=max_depth))
    fn = tree.forces_exact_leaves if exact_leaves else tree.forces
    f_ref, _ = fn(i0=lo, i1=hi, stride=stride, nthreads=oracle.max_threads())
    a_, b_ = f_gpu[lo:hi:stride], f_ref[lo:hi:stride]
    ok = np.isfinite(b_).all(axis=1)
    err = float(np.sqrt(((a_[ok] - b_[ok]) ** 2).sum() / max((b_[ok] ** 2).sum(), 1e-300)))
    return {"force_rel_rms_vs_reference_tree": err, "bar": 1e-5,
            "sample": f"every {stride}th body of [{lo}, {hi}) ({int(ok.sum())} bodies) of the timed workload; reference tree "
                      "forces from the C restatement pinned bit-for-bit to the reference (oracle/)"}


def run_ours(args):
    import torch
    import gpu_nbody_simulation_b200 as bh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    numa, numa_why = pin_to_gpu_numa_node_why(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    strong_only = args.total_bodies > 0
    n = args.total_bodies if strong_only else BODIES_PER_GPU * world
    K, W = args.steps, args.warmup
    R = args.brackets if args.brackets > 0 else (3 if args.quick else bracket_plan(K))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def morton_order(pos, vel, mass, n_):
        # Hand the bodies over in Morton order of the initial distribution (computed once with the
        # library itself, outside any timed region): a rank's contiguous index slice is then a compact
        # region, so the 64 bodies of a traversal warp stay neighbours.  The device-resident legs would not
        # need this any more (bh_snapshot re-partitions a multi-rank context: gather, full build, global
        # permutation — DESIGN 12), but the e2e leg uploads each rank's slice from the CALLER's arrays every
        # step, and there the application's order is what a rank gets.
        with bh.Simulation(n_, device=local, max_depth=args.max_depth) as tmp:
            tmp.set_bodies(pos, vel, mass)
            tmp.build_tree()
            order = tmp.sorted_order().astype(np.int64)
        return np.ascontiguousarray(pos[order]), np.ascontiguousarray(vel[order]), np.ascontiguousarray(mass[order])

    def make_sim(n_, n_ranks, counters=False, p2p=True):
        s = bh.Simulation(n_, device=local, rank=rank if n_ranks > 1 else 0, n_ranks=n_ranks, graph=not args.no_graph,
                          max_depth=args.max_depth, counters=counters, exact_leaves=args.exact_leaves)
        if n_ranks > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(bh.nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            s.attach_nccl(bytes(idt.cpu().numpy().tobytes()))   # NCCL: getters (ragged all-gather of slices)
            if p2p and not args.no_p2p:
                # per-step exchange over NVLink peer memory, fused with the kernels (csrc/peer_comm.cu)
                handles = [None] * n_ranks
                dist.all_gather_object(handles, s.comm_handle())
                s.attach_peers(handles)
                dist.barrier()      # nobody stores into a peer before every rank has opened every handle
        return s

    def timed_brackets(s, k, w, r):
        """w warm-up steps, then r brackets of EXACTLY k steps, each between barrier + synchronize fences; per bracket
        the device time from CUDA events on the library's stream, max over ranks.  Returns the list of ms."""
        barrier()
        s.step_from_snapshot(w)
        s.synchronize()
        out = []
        for _ in range(r):
            barrier()
            s.step_from_snapshot(k)
            barrier()
            out.append(max_over_ranks(s.last_step_ms()))
        return out

    pos, vel, mass = make_workload(n, args.dist)
    if world > 1 and not args.no_presort:
        pos, vel, mass = morton_order(pos, vel, mass, n)
    sim = make_sim(n, world)
    sim.set_bodies(pos, vel, mass)
    sim.snapshot()

    # ---- device-resident timing -------------------------------------------------------------------
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2(i):
        flush_buf.fill_(i & 0xFF)
        torch.cuda.synchronize()

    def timed_steps_flushed(s, k):
        tot = 0.0
        for i in range(k):
            flush_l2(i)
            s.step_from_snapshot(1)
            tot += s.last_step_ms()
        return tot

    # clocks: ONE nvidia-smi for the whole job (rank 0 samples every GPU of the job), started and
    # producing samples BEFORE the warm-up so that its start-up cannot disturb the timed brackets
    vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
    smi_ids = [vis[i] if i < len(vis) else str(i) for i in range(world)]   # nvidia-smi ignores CUDA_VISIBLE_DEVICES
    sampler = ClockSampler(",".join(smi_ids)) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    sim.step_from_snapshot(W)
    sim.synchronize()
    barrier()
    if sampler:
        sampler.mark()
    sim.reset_timers()
    bracket_ms = timed_brackets(sim, K, 0, R)
    launches = sim.timers()["kernel_launches"] // R        # kernels of this library per K-step bracket
    ms = statistics.median(bracket_ms)
    # variant with an explicit L2 flush before every step (steps timed one by one and summed); with several
    # ranks this one also charges every host-side skew between the ranks to the waiting rank
    ms_flushed = max_over_ranks(timed_steps_flushed(sim, K))
    barrier()
    clocks = sampler.stop() if sampler else None
    value = n * K / (ms * 1e-3)

    # ---- per-phase events + interaction count (second pass, direct launches, same work) -----------
    sim.set_profiling(True)
    sim.step_from_snapshot(2)
    sim.reset_timers()
    timed_steps_flushed(sim, min(K, 50))
    sim.synchronize()
    t = sim.timers()
    sim.set_profiling(False)
    phases = {k: t[k] / max(t["steps"], 1) for k in ("bounds_keys_us", "sort_us", "build_us", "traverse_us",
                                                     "exchange_us", "total_us")}
    phases_all = None
    if dist is not None:      # every rank's phases (rank skew shows up as waiting time in the phases that exchange)
        phases_all = [None] * world
        dist.all_gather_object(phases_all, {k: round(v, 1) for k, v in phases.items()})
    # interactions of THIS rank's bodies: counting variant of the traversal kernel on a second context (multi-rank:
    # NCCL exchange, one step)
    simc = make_sim(n, world, counters=True, p2p=False)
    simc.set_bodies(pos, vel, mass)
    simc.snapshot()
    simc.step_from_snapshot(1)
    simc.synchronize()
    inter_rank = simc.counters()["interactions"]
    simc.close()
    roofline = None
    own_lo, own_hi = bh.shard_range(n, world, rank)
    if rank == 0:
        peak_tf, mhz = bh.measure_fp32_peak(local)
        trav_s = phases["traverse_us"] * 1e-6
        achieved = inter_rank * FLOP_PER_INTERACTION / trav_s / 1e12
        roofline = {"bound": "fp32", "kernel": "traverse_kernel<fp32,integrate>", "achieved": achieved,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "traffic": TRAVERSE_DRAM_BYTES_NCU if (world == 1 and n == BODIES_PER_GPU) else None,
                    "traffic_source": f"dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one ncu --set full "
                                      f"capture at N=1M ({NCU_SOURCE}); algorithmic bytes = 72 B/body state + 11 MB tree = 83 MB",
                    "peak_source": f"measured here: FFMA loop, {peak_tf:.1f} TFLOP/s (implies {mhz:.0f} MHz at 128 FMA/clk/SM); "
                                   "FP32 peak is not in MEASURED_PEAKS.json (SURVEY 8d)",
                    "interactions_per_step": inter_rank, "bodies": own_hi - own_lo,
                    "flop_per_interaction": FLOP_PER_INTERACTION, "kernel_us": phases["traverse_us"],
                    "kernel_timing": "cudaEvents around the kernel on the library's stream, averaged over a second pass "
                                     "of steps with direct launches (the timed pass replays a CUDA graph); rank 0's kernel "
                                     "and rank 0's bodies",
                    "interactions_per_s": inter_rank / trav_s}
        if world == 1 and n == BODIES_PER_GPU and args.dist == "disk" and args.max_depth == 10 and not args.exact_leaves:
            sms = torch.cuda.get_device_properties(local).multi_processor_count
            slots_per_s = sms * 4 * mhz * 1e6
            roofline["issue_view"] = {"warp_instructions_per_launch": TRAVERSE_WARP_INSTRUCTIONS_NCU,
                                      "source": f"smsp__inst_executed.sum, {NCU_SOURCE}",
                                      "issue_slots_per_s": slots_per_s, "sm_count": sms, "sm_mhz_from_fma_loop": mhz,
                                      "frac_of_issue_slots": TRAVERSE_WARP_INSTRUCTIONS_NCU / (slots_per_s * trav_s)}
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            src = "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, src = 6650.0, "fallback"
        roofline["hbm_view"] = {"bytes_per_body_step": 72, "achieved_gbs": 72.0 * (n // world) / (ms / K * 1e-3) / 1e9,
                                "peak_gbs": hbm_peak, "peak_source": src,
                                "note": "per GPU; the step is FP32-issue bound, not HBM bound (SURVEY 8d)"}

    # ---- accuracy of the timed configuration (BASELINE.json's metric carries "force rel-RMS error vs ref") ---------
    # forces of the default (FP32) traversal; multi-rank: the getter gathers every rank's slice, rank 0 checks ITS slice
    accuracy = None
    sim.restore()
    sim.build_tree()
    sim.compute_forces()
    f_gpu = sim.forces()              # collective on a multi-rank context
    if rank == 0 and not args.no_cpu_baseline:
        try:
            import oracle
            accuracy = sampled_accuracy(bh, oracle, f_gpu, pos, mass, args.max_depth, own_lo, own_hi,
                                        exact_leaves=args.exact_leaves)
        except Exception as e:   # never worth losing the line for
            accuracy = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    del f_gpu

    # ---- end to end through the C-ABI with host buffers --------------------------------------------
    e2e = None
    if not args.no_e2e:
        hp = torch.from_numpy(pos).pin_memory()
        hv = torch.from_numpy(vel).pin_memory()
        hm = torch.from_numpy(mass).pin_memory()
        hout = torch.empty((n, 2), dtype=torch.float64).pin_memory()
        for _ in range(max(1, min(W, 3))):
            sim.step_host(hp, hv, hm, hout)
        e2e_runs = []
        for _ in range(1 if args.quick else max(3, min(R, 10))):
            barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                # bh_step_host: H2D of this step's inputs (pinned), one step, D2H of its result; synchronous
                sim.step_host(hp, hv, hm, hout)
            barrier()
            e2e_runs.append(max_over_ranks(time.perf_counter() - t0))
        e2e_s = statistics.median(e2e_runs)
        per_rank = own_hi - own_lo
        e2e = {"value": n * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(40 * n), "d2h_bytes_per_step": int(16 * n),
               "ms_per_step": e2e_s / K * 1e3, "brackets": len(e2e_runs),
               "note": f"each rank moves its own slice ({per_rank} bodies: {40 * per_rank} B up, {16 * per_rank} B down) over its own PCIe link"}
        del hp, hv, hm, hout

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not strong_only:
        _, cpu_baseline = cpu_reference_run(BODIES_PER_GPU, 3, 1, budget_s=30.0)
        cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}
        try:
            cpu_baseline["port_all_cores"] = cpu_port_all_cores(pos, mass)
        except Exception as e:   # an extra, never worth losing the line for
            cpu_baseline["port_all_cores"] = {"unavailable": str(e)[:200]}

    gpu_baseline = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline and not strong_only and args.dist == "disk" and args.max_depth == 10:
        gpu_baseline = reference_gpu_run(pos, vel, mass, local)

    # ---- config 5: direct all-pairs kernel at 262 144 bodies (single GPU) -------------------------------------
    direct = None
    if rank == 0 and world == 1 and not args.no_direct and not strong_only:
        try:
            dn = DIRECT_N
            with bh.Simulation(dn, device=local) as sd:
                sd.set_bodies(pos[:dn], vel[:dn], mass[:dn])
                sd.direct_forces(want_output=False)
                times = [sd.direct_forces(want_output=False)[1] for _ in range(5)]
                f_d, _ = sd.direct_forces()
            dms = statistics.median(times)
            pairs = float(dn) * dn
            peak_tf = roofline["peak"] if roofline else None
            direct = {"n_bodies": dn, "ms": dms, "pairs_per_s": pairs / (dms * 1e-3),
                      "tflops_at_20_flop_per_pair": pairs * 20.0 / (dms * 1e-3) / 1e12,
                      "frac_of_fp32_peak": (pairs * 20.0 / (dms * 1e-3) / 1e12 / peak_tf) if peak_tf else None,
                      "formula": "main_approach_1.cpp:53-75 (G m_i m_j d / (d^2 d), i != j), FP32 arithmetic on single-float "
                                 "coordinates (FP64 re-centring before the conversion), shared-memory tiles"}
            if not args.no_cpu_baseline:
                import oracle
                sel = np.linspace(0, dn - 1, 64).astype(np.int64)
                ref = np.stack([oracle.direct_forces(pos[:dn], mass[:dn], i0=int(i), i1=int(i) + 1,
                                                     nthreads=oracle.max_threads())[int(i)] for i in sel])
                direct["force_rel_rms_vs_oracle_direct_sum"] = float(
                    np.sqrt(((f_d[sel] - ref) ** 2).sum() / (ref ** 2).sum()))
                direct["sample"] = "64 bodies x all 262 144 partners, FP64 direct sum of the pinned oracle"
        except Exception as e:
            direct = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    sim.close()
    del flush_buf
    torch.cuda.empty_cache()

    # ---- strong scaling at 16M bodies (north_star's multi-GPU target; BASELINE config 4) ------------------------
    strong = None
    if world > 1 and not strong_only and not args.no_strong:
        ns = STRONG_TOTAL
        Ks = max(3, min(K, 20))
        Rs = 5
        sp, sv, sm_ = make_workload(ns, "disk")
        sp, sv, sm_ = morton_order(sp, sv, sm_, ns)
        ss = make_sim(ns, world)
        ss.set_bodies(sp, sv, sm_)
        ss.snapshot()
        s_ms = statistics.median(timed_brackets(ss, Ks, 3, Rs)) / Ks
        ss.set_profiling(True)
        ss.step_from_snapshot(2)
        ss.reset_timers()
        ss.step_from_snapshot(Ks)
        ss.synchronize()
        ts = ss.timers()
        ss.set_profiling(False)
        s_phases = {k: ts[k] / max(ts["steps"], 1) for k in ("bounds_keys_us", "sort_us", "build_us", "traverse_us",
                                                             "exchange_us", "total_us")}
        ss.restore()
        ss.build_tree()
        ss.compute_forces()
        fs = ss.forces()
        ss.close()
        barrier()
        one_ms, s_acc = None, None
        if rank == 0:
            lo0, hi0 = bh.shard_range(ns, world, 0)
            if not args.no_cpu_baseline:
                try:
                    import oracle
                    s_acc = sampled_accuracy(bh, oracle, fs, sp, sm_, args.max_depth, lo0, hi0, nsample=2048)
                except Exception as e:
                    s_acc = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
            with bh.Simulation(ns, device=local, max_depth=args.max_depth, graph=not args.no_graph) as s1:
                s1.set_bodies(sp, sv, sm_)
                s1.snapshot()
                s1.step_from_snapshot(3)
                s1.synchronize()
                runs = []
                for _ in range(Rs):
                    torch.cuda.synchronize()
                    s1.step_from_snapshot(Ks)
                    runs.append(s1.last_step_ms())
                one_ms = statistics.median(runs) / Ks
        barrier()
        if rank == 0:
            strong = {"total_bodies": ns, "n_gpus": world, "steps_per_bracket": Ks, "brackets": Rs,
                      "ms_per_step": s_ms, "value": ns / (s_ms * 1e-3), "unit": UNIT,
                      "ms_per_step_1gpu_same_job": one_ms, "speedup_vs_1gpu": one_ms / s_ms,
                      "target": "north_star: >= 6x at 16M bodies on 8 GPUs",
                      "phases_us_rank0": s_phases, "accuracy": s_acc,
                      "workload": workload_string(ns, world)}
        del sp, sv, sm_, fs

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if strong_only else "weak",
                "vs_baseline": None,
                "dtype": "f32 (double-float displacement; FP64 state, tree and integrator)", "data": "synthetic",
                "brackets": {"count": R, "steps_each": K, "ms_median": ms, "ms_min": min(bracket_ms), "ms_max": max(bracket_ms),
                             "note": "value = N K / median over the brackets; every bracket is EXACTLY K steps between "
                                     "barrier + synchronize fences, device time, max over ranks"},
                "config": {"workload": workload_string(n, world, args.dist, args.max_depth), "n_bodies": n,
                           "exact_leaves": bool(args.exact_leaves),
                           "l2": "inputs larger than L2: the per-rank working set that every step reads and rewrites is "
                                 "~150 MB at 1M bodies per GPU (FP64 state + snapshot 116 MB, sort buffers 16 MB, tree 23 MB) "
                                 "vs 126 MB of L2; value_l2_flushed repeats K steps with 512 MB written before each "
                                 "step (steps timed one by one and summed; with several ranks the un-timed flush also "
                                 "absorbs rank skew, so the bracketed value is the one to quote)",
                           "parallelism": (f"morton-shard x{world}: bodies handed over in Morton order, contiguous index "
                                           f"slice per rank, sharded build, " + ("2 NCCL all-reduces per step" if args.no_p2p else
                                           "box + cell-sum exchange by NVLink peer stores fused with the kernels")) if world > 1
                           else "single GPU", "host_numa_node": numa, "host_numa_note": numa_why},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "accuracy": accuracy, "cpu_baseline": cpu_baseline, "gpu_baseline": gpu_baseline,
                "phases_us": phases, "phases_us_all_ranks": phases_all, "strong": strong, "direct": direct,
                "value_l2_flushed": n * K / (ms_flushed * 1e-3)}
        print(json.dumps(line), flush=True)
        if args.reference_lines:
            # the reference program's two stdout lines (project.cu:1097, :1102) for the K timed steps, so that
            # scripts/gpu_scaling_script.sh can feed the reference's plot scripts
            par_us = (phases["traverse_us"] if phases else ms / K * 1e3) * K
            print(f"GPU total computation took {int(round(ms))} milliseconds. "
                  f"GPU parallel computation took {int(round(par_us))} microseconds.", flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--brackets", type=int, default=0, help="number of K-step brackets (default: ~1000 steps in total, 5..50)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference project.cu run on this GPU")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the 16M-body strong-scaling object")
    ap.add_argument("--no-direct", action="store_true", help="N = 1: skip the direct all-pairs object (config 5)")
    ap.add_argument("--dist", choices=["disk", "plummer", "square"], default="disk",
                    help="synthetic distribution (BASELINE config 2: disk; config 3: plummer)")
    ap.add_argument("--max-depth", type=int, default=10,
                    help="QUADTREE_MAX_DEPTH (reference: 10); BASELINE config 3 (clustered Plummer) raises it, up to 13")
    ap.add_argument("--exact-leaves", action="store_true",
                    help="BH_FLAG_EXACT_LEAVES (extension); N > 1: every step all-gathers the positions, every rank builds the full tree")
    ap.add_argument("--no-graph", action="store_true", help="direct kernel launches instead of CUDA-graph replay")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL all-reduces instead of the peer-memory exchange")
    ap.add_argument("--total-bodies", type=int, default=0,
                    help="strong scaling: total body count over all GPUs (default: weak scaling, 1M per GPU)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (very large N: pinned host memory)")
    ap.add_argument("--no-presort", action="store_true",
                    help="N > 1: hand the bodies over in the generator's (random) order and let the engine re-partition")
    ap.add_argument("--quick", action="store_true",
                    help="headline brackets + phases only (no e2e, accuracy, baselines, strong, direct): scaling scripts")
    ap.add_argument("--reference-lines", action="store_true",
                    help="also print the reference program's two timing lines (for scripts/gpu_scaling_script.sh)")
    args = ap.parse_args()
    if args.quick:
        args.no_cpu_baseline = args.no_gpu_baseline = args.no_strong = args.no_direct = True
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
