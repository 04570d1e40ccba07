#!/usr/bin/env python
"""Benchmark of the Barnes-Hut hot path (BASELINE.json: body·steps/s, theta = 0.5, 2-D).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] — 1 000 000 bodies, uniform disk R = 0.1, seed
12345, the reference's constants (G, dt = 1, theta = 0.5, depth cap 10).  A *step* is one pass of
the whole hot path: bounds -> cell keys -> radix sort -> tree + COM -> traversal -> integrate.
Because the reference's own physics flings bodies away after ONE step and the tree collapses to
a few hundred nodes, every step starts from the initial distribution (a device
snapshot that the step reads out of place; nothing is skipped: bounds, keys, sort, tree, traversal
and integrator all run every step) — i.e. every timed step does the full-size, non-degenerate work
of the reference's step 0.  At N > 1 GPUs the body count grows with N (weak scaling, 1M bodies per
GPU, contiguous Morton slices, sharded tree build; the ranks exchange the bounding box and the
per-cell sums over NVLink peer memory, not the bodies).

`value`  : device-resident throughput, inputs in HBM: W warm-up steps, then EXACTLY K steps between two
           barrier + torch.cuda.synchronize() brackets, device time from CUDA events on the library's
           stream, max over ranks.  The per-rank working set (~150 MB) exceeds the 126 MB L2;
           `value_l2_flushed` repeats the K steps with an explicit L2 flush before each one.
`e2e`    : same step through the C-ABI with HOST buffers: pinned H2D of positions, velocities and
           masses + step + D2H of positions every step, wall clock around the synchronous calls.
`roofline`: traversal kernel, 20 flop per accepted interaction against the FP32
           FMA peak measured by a register-resident FMA loop on the same device; `issue_view` = the same
           kernel against the SM's issue slots (what actually binds it).
`cpu_baseline`: the reference's OWN CPU functions (oracle/_ref, 1 thread: it has no threading) on the same
           bodies (+ `port_all_cores`: the oracle port with OpenMP over bodies, labelled as a port).
`gpu_baseline`: the reference's OWN GPU program (unmodified project.cu, sm_100a) on the same GPU, its two timers.
`accuracy`: force rel-RMS of the timed configuration against the reference tree forces (sampled bodies).
`--impl reference`: the CPU reference arm alone, same JSON line shape.
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
SEED = 12345
FLOP_PER_INTERACTION = 20.0
TRAVERSE_DRAM_BYTES_NCU = 74_801_152 + 37_312_768   # profiles/r01_traverse_v8_pair_ncu_summary.txt
TRAVERSE_WARP_INSTRUCTIONS_NCU = 298_806_710          # same capture: smsp__inst_executed.sum at N = 1M uniform disk
METRIC = "body_steps_per_s"
UNIT = "body·steps/s"


def make_workload(n, dist="disk"):
    from gpu_nbody_simulation_b200 import initial_conditions as ic
    # values pass through the reference writers' "%.6g" text format (round6) up to 2M bodies; above that
    # the string round trip alone takes minutes per rank, so the raw FP64 draws are used
    gen = {"disk": ic.uniform_disk, "plummer": ic.plummer_2d, "square": ic.uniform_square}[dist]
    return gen(n, seed=SEED, round6=n <= 2_000_000)


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
        30-60 % (gpurun_out/final_weak_g8.json vs bench_weak_g8.log).  Polling afterwards is cheap."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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


def pin_to_gpu_numa_node(dev):
    """Binds this process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that the pinned
    host buffers of the e2e leg are allocated next to the GPU's PCIe root.  Best effort: returns the
    node number or None (no sysfs entry, single node, restricted cpuset)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(dev)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def cpu_reference_run(n_bodies, steps, warmup, budget_s=150.0):
    """Times the reference's OWN CPU path (oracle/_ref, unmodified project.cu functions, 1 thread —
    the reference has no threading) or, if it was not built, the oracle port.  Every step restarts
    from the initial distribution, like the GPU arm.  Returns (value, info)."""
    import oracle
    est_per_step = {1_000_000: 10.5, 262_144: 2.6, 65_536: 0.6}
    total = steps + warmup
    choice = None
    for n in (1_000_000, 262_144, 65_536):
        if n <= n_bodies and oracle.ref_available(n) and est_per_step[n] * total <= budget_s:
            choice = n
            break
    pos, vel, mass = make_workload(1_000_000)
    if choice is not None:
        p, v, m = pos[:choice], vel[:choice], mass[:choice]
        _, tim = oracle.run_ref(p, v, m, steps=total, dump="", keep_dump=False, reset_each_step=True)
        timed = tim[warmup:]
        secs = sum(t["build_us"] + t["force_us"] + t["update_us"] for t in timed) * 1e-6
        value = choice * len(timed) / secs
        info = {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": (f"{len(timed)} full steps (buildTree + computeForces + update*) of the reference's own CPU "
                           f"functions on the first {choice} bodies of the 1M uniform disk, restarted from the initial "
                           f"distribution every step; build {sum(t['build_us'] for t in timed) / len(timed) / 1e3:.0f} ms, "
                           f"force {sum(t['force_us'] for t in timed) / len(timed) / 1e3:.0f} ms per step"),
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
    value = 1_000_000 / secs
    return value, {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"oracle port: full 1M tree build + forces for every {stride}th body, extrapolated",
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
    value, info = cpu_reference_run(1_000_000, args.steps, args.warmup)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "uniform disk N=1M, theta=0.5, reference constants, every step from the initial "
                                   "distribution (CPU reference path, see cpu_baseline.sample)"},
            "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    numa = pin_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    strong = args.total_bodies > 0
    n = args.total_bodies if strong else BODIES_PER_GPU * world
    pos, vel, mass = make_workload(n, args.dist)
    K, W = args.steps, args.warmup

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

    if world > 1:
        # Hand the bodies over in Morton order of the initial distribution (computed once with the
        # library itself, outside any timed region): a rank's contiguous index slice is then a compact
        # region, so the 64 bodies of a traversal warp stay neighbours.  Body order is arbitrary for a
        # synthetic workload; a multi-GPU application keeps its bodies in this order permanently.
        with bh.Simulation(n, device=local, max_depth=args.max_depth) as tmp:
            tmp.set_bodies(pos, vel, mass)
            tmp.build_tree()
            order = tmp.sorted_order().astype(np.int64)
        pos, vel, mass = np.ascontiguousarray(pos[order]), np.ascontiguousarray(vel[order]), np.ascontiguousarray(mass[order])
    sim = bh.Simulation(n, device=local, rank=rank, n_ranks=world, graph=not args.no_graph, max_depth=args.max_depth)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(bh.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        sim.attach_nccl(bytes(idt.cpu().numpy().tobytes()))   # NCCL: getters (ragged all-gather of slices)
        if not args.no_p2p:
            # per-step exchange over NVLink peer memory, fused with the kernels (csrc/peer_comm.cu)
            handles = [None] * world
            dist.all_gather_object(handles, sim.comm_handle())
            sim.attach_peers(handles)
    sim.set_bodies(pos, vel, mass)
    sim.snapshot()

    # ---- device-resident timing -------------------------------------------------------------------
    # L2 is flushed (512 MB written) before every timed step; each step is timed on its own with
    # CUDA events on the library's stream and the K durations are summed.
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2(i):
        flush_buf.fill_(i & 0xFF)
        torch.cuda.synchronize()

    def timed_steps(k):
        tot = 0.0
        for i in range(k):
            flush_l2(i)
            sim.step_from_snapshot(1)
            tot += sim.last_step_ms()
        return tot

    # clocks: ONE nvidia-smi for the whole job (rank 0 samples every GPU of the job), started and
    # producing samples BEFORE the warm-up so that its start-up cannot disturb the timed bracket
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
    # headline: EXACTLY K steps between two barrier + synchronize brackets, device time from CUDA events on
    # the library's stream (recorded around the K steps), max over ranks
    sim.reset_timers()
    sim.step_from_snapshot(K)
    barrier()
    ms = max_over_ranks(sim.last_step_ms())
    launches = sim.timers()["kernel_launches"]
    # variant with an explicit L2 flush before every step (steps timed one by one and summed); with several
    # ranks this one also charges every host-side skew between the ranks to the waiting rank
    ms_flushed = max_over_ranks(timed_steps(K))
    barrier()
    clocks = sampler.stop() if sampler else None
    value = n * K / (ms * 1e-3)

    # ---- per-phase events + interaction count (second pass, direct launches, same work) -----------
    phases, inter_per_step, roofline = None, None, None
    if world > 1:      # collective: every rank runs the profiled pass, rank 0 reports its phases
        sim.set_profiling(True)
        sim.step_from_snapshot(2)
        sim.reset_timers()
        timed_steps(K)
        sim.synchronize()
        t = sim.timers()
        sim.set_profiling(False)
        phases = {k: t[k] / max(t["steps"], 1) for k in ("bounds_keys_us", "sort_us", "build_us", "traverse_us",
                                                         "exchange_us", "total_us")}
    simc = bh.Simulation(n, device=local, rank=rank, n_ranks=1, counters=True, max_depth=args.max_depth) if world == 1 else None
    if rank == 0 and simc is not None:
        simc.set_bodies(pos, vel, mass)
        simc.snapshot()
        simc.step_from_snapshot(1)
        simc.synchronize()
        cnt = simc.counters()
        inter_per_step = cnt["interactions"]
        simc.close()
        sim.set_profiling(True)
        sim.step_from_snapshot(2)
        sim.reset_timers()
        timed_steps(K)
        sim.synchronize()
        t = sim.timers()
        sim.set_profiling(False)
        phases = {k: t[k] / max(t["steps"], 1) for k in ("bounds_keys_us", "sort_us", "build_us", "traverse_us",
                                                         "exchange_us", "total_us")}
        peak_tf, mhz = bh.measure_fp32_peak(local)
        trav_s = phases["traverse_us"] * 1e-6
        achieved = inter_per_step * FLOP_PER_INTERACTION / trav_s / 1e12
        roofline = {"bound": "fp32", "kernel": "traverse_kernel<fp32,integrate>", "achieved": achieved,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "traffic": TRAVERSE_DRAM_BYTES_NCU,
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one ncu --set full "
                                      "capture at N=1M (profiles/r01_traverse_v8_pair_ncu_summary.txt); algorithmic "
                                      "bytes = 72 B/body state + 11 MB tree = 83 MB",
                    "peak_source": f"measured here: FFMA loop, {peak_tf:.1f} TFLOP/s (implies {mhz:.0f} MHz at 128 FMA/clk/SM); "
                                   "FP32 peak is not in MEASURED_PEAKS.json (SURVEY 8d)",
                    "interactions_per_step": inter_per_step, "flop_per_interaction": FLOP_PER_INTERACTION,
                    "kernel_us": phases["traverse_us"],
                    "kernel_timing": "cudaEvents around the kernel on the library's stream, averaged over a second pass "
                                     "of the same K steps with direct launches (the timed pass replays a CUDA graph)",
                    "interactions_per_s": inter_per_step / trav_s}
        if n == 1_000_000 and args.dist == "disk" and args.max_depth == 10:
            # issue-slot view of the same kernel: the 20-flop convention counts accepted interactions, the hardware
            # spends ~18.4 warp instructions per (body, node) EVALUATION; instructions per launch are a
            # property of kernel + workload, measured once with ncu
            sms = torch.cuda.get_device_properties(local).multi_processor_count
            slots_per_s = sms * 4 * mhz * 1e6
            roofline["issue_view"] = {"warp_instructions_per_launch": TRAVERSE_WARP_INSTRUCTIONS_NCU,
                                      "source": "smsp__inst_executed.sum, profiles/r01_traverse_v8_pair_ncu_summary.txt",
                                      "issue_slots_per_s": slots_per_s, "sm_count": sms, "sm_mhz_from_fma_loop": mhz,
                                      "frac_of_issue_slots": TRAVERSE_WARP_INSTRUCTIONS_NCU / (slots_per_s * trav_s)}
        # whole-step HBM view: mandatory body traffic (72 B/body FP64 state) vs measured copy peak
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            src = "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, src = 6650.0, "fallback"
        roofline["hbm_view"] = {"bytes_per_body_step": 72, "achieved_gbs": 72.0 * n / (ms / K * 1e-3) / 1e9,
                                "peak_gbs": hbm_peak, "peak_source": src,
                                "note": "the step is FP32-issue bound, not HBM bound (SURVEY 8d)"}

    # ---- end to end through the C-ABI with host buffers --------------------------------------------
    hp = torch.from_numpy(pos).pin_memory()
    hv = torch.from_numpy(vel).pin_memory()
    hm = torch.from_numpy(mass).pin_memory()
    hout = torch.empty((n, 2), dtype=torch.float64).pin_memory()
    for _ in range(max(1, min(W, 3))):
        sim.step_host(hp, hv, hm, hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        # bh_step_host: H2D of this step's inputs (pinned), one step, D2H of its result; synchronous
        sim.step_host(hp, hv, hm, hout)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": n * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(40 * n), "d2h_bytes_per_step": int(16 * n),
           "ms_per_step": e2e_s / K * 1e3}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, cpu_baseline = cpu_reference_run(1_000_000, 1, 0, budget_s=30.0)
        cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}
        try:
            cpu_baseline["port_all_cores"] = cpu_port_all_cores(pos, mass)
        except Exception as e:   # an extra, never worth losing the line for
            cpu_baseline["port_all_cores"] = {"unavailable": str(e)[:200]}

    # accuracy of the timed configuration (BASELINE.json's metric carries "force rel-RMS error vs ref"): forces of
    # the default (FP32) traversal against the reference tree forces from the pinned oracle, on a body sample
    accuracy = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            with bh.Simulation(n, device=local, max_depth=args.max_depth) as sa:
                sa.set_bodies(pos, vel, mass)
                sa.build_tree()
                sa.compute_forces()
                f_gpu = sa.forces()
            stride = max(1, n // 4096)
            tree = oracle.Tree(pos, mass, oracle.default_params(max_depth=args.max_depth))
            f_ref, _ = tree.forces(stride=stride, nthreads=oracle.max_threads())
            a_, b_ = f_gpu[::stride], f_ref[::stride]
            ok = np.isfinite(b_).all(axis=1)
            err = float(np.sqrt(((a_[ok] - b_[ok]) ** 2).sum() / max((b_[ok] ** 2).sum(), 1e-300)))
            accuracy = {"force_rel_rms_vs_reference_tree": err, "bar": 1e-5,
                        "sample": f"every {stride}th body ({int(ok.sum())} bodies) of the timed workload; reference tree "
                                  "forces from the C restatement pinned bit-for-bit to the reference (oracle/)"}
        except Exception as e:   # never worth losing the line for
            accuracy = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    gpu_baseline = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline and not strong and args.dist == "disk" and args.max_depth == 10:
        gpu_baseline = reference_gpu_run(pos, vel, mass, local)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None,
                "dtype": "f32 (double-float displacement; FP64 state, tree and integrator)", "data": "synthetic",
                "config": {"workload": f"{'uniform disk' if args.dist == 'disk' else args.dist} N={n} ({n // world} per GPU), R=0.1, seed {SEED}, theta=0.5, "
                                       f"G=6.67e-11, dt=1, depth cap {args.max_depth}; every step restarts from the device-resident snapshot of the "
                                       "initial distribution (out of place: the step reads the snapshot and writes the live "
                                       "state, so every timed step is the full-size work of the reference's step 0)",
                           "l2": "inputs larger than L2: the per-rank working set that every step reads and rewrites is "
                                 "~150 MB at 1M bodies per GPU (FP64 state + snapshot 116 MB, sort buffers 16 MB, tree 23 MB) "
                                 "vs 126 MB of L2; value_l2_flushed repeats the K steps with 512 MB written before each "
                                 "step (steps timed one by one and summed; with several ranks the un-timed flush also "
                                 "absorbs rank skew, so the bracketed value is the one to quote)",
                           "parallelism": (f"morton-shard x{world}: bodies handed over in Morton order, contiguous index "
                                           f"slice per rank, sharded build, " + ("2 NCCL all-reduces per step" if args.no_p2p else
                                           "box + cell-sum exchange by NVLink peer stores fused with the kernels")) if world > 1
                           else "single GPU", "host_numa_node": numa},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "accuracy": accuracy, "cpu_baseline": cpu_baseline, "gpu_baseline": gpu_baseline,
                "phases_us": phases,
                "value_l2_flushed": n * K / (ms_flushed * 1e-3)}
        print(json.dumps(line), flush=True)
        if args.reference_lines:
            # the reference program's two stdout lines (project.cu:1097, :1102) for the K timed steps, so that
            # scripts/gpu_scaling_script.sh can feed the reference's plot scripts
            par_us = (phases["traverse_us"] if phases else ms / K * 1e3) * K
            print(f"GPU total computation took {int(round(ms))} milliseconds. "
                  f"GPU parallel computation took {int(round(par_us))} microseconds.", flush=True)
    sim.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference project.cu run on this GPU")
    ap.add_argument("--dist", choices=["disk", "plummer", "square"], default="disk",
                    help="synthetic distribution (BASELINE config 2: disk; config 3: plummer)")
    ap.add_argument("--max-depth", type=int, default=10,
                    help="QUADTREE_MAX_DEPTH (reference: 10); BASELINE config 3 (clustered Plummer) raises it, up to 13")
    ap.add_argument("--no-graph", action="store_true", help="direct kernel launches instead of CUDA-graph replay")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL all-reduces instead of the peer-memory exchange")
    ap.add_argument("--total-bodies", type=int, default=0,
                    help="strong scaling: total body count over all GPUs (default: weak scaling, 1M per GPU)")
    ap.add_argument("--reference-lines", action="store_true",
                    help="also print the reference program's two timing lines (for scripts/gpu_scaling_script.sh)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
