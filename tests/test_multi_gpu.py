"""Multi-GPU path (SURVEY 8e / DESIGN 7): contiguous index slices, sharded build, exchange of box + cell sums.

* `gpu` test: 2 ranks over NCCL on a box with >= 2 GPUs (skipped on the 1-GPU round-end box);
  tests/multi_gpu_check.py compares against a single-GPU context and the CPU oracle.
* CPU tests (gloo, world_size 2, run everywhere): the host-side protocol — shard ranges from the
  C-ABI (bh_shard_range), owned-slice update, all-gather — with the oracle standing in for the
  device kernels, against the single-rank oracle trajectory.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_two_gpus_match_single_gpu_and_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_check.py"), "--bodies", "100001",
           "--steps", "2"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_CHECK PASS" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def _gloo_worker(rank, world, port, n, steps, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import gpu_nbody_simulation_b200 as bh
    import oracle
    from gpu_nbody_simulation_b200 import initial_conditions as ic
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos, vel, mass = ic.uniform_disk(n, seed=99, round6=False)
    par = oracle.default_params(G=6.67e-17)
    lo, hi = bh.shard_range(n, world, rank)               # same integer logic the CUDA library uses
    sizes = [bh.shard_range(n, world, r) for r in range(world)]
    p, v = pos.copy(), vel.copy()
    for _ in range(steps):
        tree = oracle.Tree(p, mass, par)                   # every rank rebuilds the whole tree
        f, _ = tree.forces(i0=lo, i1=hi)                   # forces for the owned slice only
        a, vv, pp = oracle.update(f[lo:hi], mass[lo:hi], v[lo:hi], p[lo:hi], par.dt)
        v[lo:hi] = vv
        # ragged slices: one broadcast per owner, exactly what exchange_slices() in csrc/api.cu groups
        parts = []
        for r, (l, h) in enumerate(sizes):
            t = torch.from_numpy(np.ascontiguousarray(pp)) if r == rank else torch.empty((h - l, 2), dtype=torch.float64)
            dist.broadcast(t, src=r)
            parts.append(t.numpy())
        p = np.concatenate(parts, axis=0)
    np.save(os.path.join(out_dir, f"pos_rank{rank}.npy"), p)
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 1001])
def test_gloo_world2_shard_and_allgather_protocol(tmp_path, n):
    import torch.multiprocessing as mp
    import oracle
    from gpu_nbody_simulation_b200 import initial_conditions as ic
    steps, world = 3, 2
    port = _free_port()
    mp.spawn(_gloo_worker, args=(world, port, n, steps, str(tmp_path)), nprocs=world, join=True)
    pos, vel, mass = ic.uniform_disk(n, seed=99, round6=False)
    par = oracle.default_params(G=6.67e-17)
    p, v = pos.copy(), vel.copy()
    for _ in range(steps):
        r = oracle.step(p, v, mass, par)
        p, v = r["pos"], r["vel"]
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, f"pos_rank{rank}.npy"))
        assert np.array_equal(got, p), f"rank {rank}: sharded trajectory differs from the single-rank one"


# ------------------------------------------------------------------------------------------------
# The protocol the CUDA library actually runs: every rank keys and sums ONLY its own slice,
# two reductions (box: min of 4 doubles; per finest cell: count, m, m x, m y) make the tree global, the
# level pass runs redundantly on the reduced sums.  Here with gloo and numpy standing in for NVLink and
# the kernels: the tree every rank ends up with must be the oracle's tree of ALL bodies.
# ------------------------------------------------------------------------------------------------
def _pyramid_from_sums(cnt, m, mx, my, finest):
    """Level pass on reduced finest-cell sums -> per level (count, mass, comx, comy), Morton order."""
    with np.errstate(invalid="ignore", divide="ignore"):
        cx = np.where(m > 0, mx / m, 0.0)
        cy = np.where(m > 0, my / m, 0.0)
    levels = {finest: (cnt.astype(np.int64), m, cx, cy)}
    for l in range(finest - 1, -1, -1):
        c, mm, x, y = levels[l + 1]
        c4, m4, x4, y4 = (a.reshape(-1, 4) for a in (c, mm, x, y))
        pc = c4.sum(axis=1)
        pm = np.zeros(len(pc)); sx = np.zeros(len(pc)); sy = np.zeros(len(pc))
        for q in range(4):                                   # children 0..3, sums start from 0.0 (project.cu:480-495)
            pm = pm + m4[:, q]; sx = sx + m4[:, q] * x4[:, q]; sy = sy + m4[:, q] * y4[:, q]
        with np.errstate(invalid="ignore", divide="ignore"):
            px = np.where(pm > 0, sx / pm, 0.0); py = np.where(pm > 0, sy / pm, 0.0)
        single = pc == 1                                     # a lone body's cell copies its only non-empty child
        which = np.argmax(c4 > 0, axis=1)
        rows = np.arange(len(pc))
        pm = np.where(single, m4[rows, which], pm)
        px = np.where(single, x4[rows, which], px); py = np.where(single, y4[rows, which], py)
        levels[l] = (pc, pm, px, py)
    return levels


def _canonical_from_pyramid(levels, finest):
    """DFS pre-order (children 0..3) over { cell : every proper ancestor holds >= 2 bodies } -> rows
    [depth, mass, comx, comy, internal]."""
    rows = []

    def rec(l, code):
        c, m, x, y = (a[code] for a in levels[l])
        internal = c >= 2 and l < finest
        rows.append((l, m, x, y, 1.0 if internal else 0.0))
        if internal:
            for q in range(4):
                rec(l + 1, 4 * code + q)
    rec(0, 0)
    return np.array(rows)


def _sharded_build_worker(rank, world, port, n, max_depth, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import gpu_nbody_simulation_b200 as bh
    import oracle
    from gpu_nbody_simulation_b200 import initial_conditions as ic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos, vel, mass = ic.plummer_2d(n, seed=7, round6=False)
    lo, hi = bh.shard_range(n, world, rank)
    own_p, own_m = pos[lo:hi], mass[lo:hi]
    finest = max_depth - 1
    # 1. box: min over ranks of (xmin, -xmax, ymin, -ymax), then the reference's padding rule on every rank
    raw = torch.tensor([own_p[:, 0].min(), -own_p[:, 0].max(), own_p[:, 1].min(), -own_p[:, 1].max()], dtype=torch.float64)
    dist.all_reduce(raw, op=dist.ReduceOp.MIN)
    xmin, nxmax, ymin, nymax = raw.tolist()
    bounds = oracle.root_bounds(np.array([[xmin, ymin], [-nxmax, -nymax]]))
    # 2. own keys, own partial sums per finest cell
    keys = oracle.body_keys(own_p, bounds, max_depth).astype(np.int64)
    nc = 4 ** finest
    sums = np.stack([np.bincount(keys, minlength=nc).astype(np.float64),
                     np.bincount(keys, weights=own_m, minlength=nc),
                     np.bincount(keys, weights=own_m * own_p[:, 0], minlength=nc),
                     np.bincount(keys, weights=own_m * own_p[:, 1], minlength=nc)])
    t = torch.from_numpy(sums)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)               # 3. ONE reduction of 4 doubles per finest cell
    sums = t.numpy()
    levels = _pyramid_from_sums(sums[0], sums[1], sums[2], sums[3], finest)   # 4. redundant level pass
    np.save(os.path.join(out_dir, f"tree_rank{rank}.npy"), _canonical_from_pyramid(levels, finest))
    np.save(os.path.join(out_dir, f"bounds_rank{rank}.npy"), bounds)
    dist.destroy_process_group()


@pytest.mark.parametrize("n,max_depth", [(3000, 6), (2001, 8)])
def test_gloo_world2_sharded_build_gives_the_global_tree(tmp_path, n, max_depth):
    import torch.multiprocessing as mp
    import oracle
    from gpu_nbody_simulation_b200 import initial_conditions as ic
    world = 2
    mp.spawn(_sharded_build_worker, args=(world, _free_port(), n, max_depth, str(tmp_path)), nprocs=world, join=True)
    pos, vel, mass = ic.plummer_2d(n, seed=7, round6=False)
    par = oracle.default_params(max_depth=max_depth)
    tree = oracle.Tree(pos, mass, par)
    want = tree.canonical()            # [depth, xmin, xmax, ymin, ymax, mass, comx, comy, occupant, internal]
    got0 = np.load(os.path.join(tmp_path, "tree_rank0.npy"))
    got1 = np.load(os.path.join(tmp_path, "tree_rank1.npy"))
    assert np.array_equal(got0, got1), "all ranks must hold the identical tree (same reduced sums, same arithmetic)"
    assert np.array_equal(np.load(os.path.join(tmp_path, "bounds_rank0.npy")), oracle.root_bounds(pos)), "box bit-exact"
    assert got0.shape[0] == want.shape[0] == tree.size, "topology: same node count"
    assert np.array_equal(got0[:, 0], want[:, 0]) and np.array_equal(got0[:, 4], want[:, 9]), "same pre-order shape"
    # values: sums-then-divide instead of the sequential running average -> ~1e-16 relative, not bit-exact
    assert np.allclose(got0[:, 1], want[:, 5], rtol=1e-13, atol=0)
    nz = want[:, 5] > 0
    assert np.allclose(got0[nz, 2], want[nz, 6], rtol=1e-12, atol=1e-18)
    assert np.allclose(got0[nz, 3], want[nz, 7], rtol=1e-12, atol=1e-18)


# ------------------------------------------------------------------------------------------------
# Re-partitioning: every rank gathers the slices, computes the SAME keys / stable order / permutation and
# applies it to its full-size arrays; the fixed index slices are then contiguous Morton ranges and the getters
# (gather, then un-permute) still return the caller's order.  gloo + numpy model of csrc/api.cu: enqueue_reorder.
# ------------------------------------------------------------------------------------------------
def _repartition_worker(rank, world, port, n, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import gpu_nbody_simulation_b200 as bh
    import oracle
    from gpu_nbody_simulation_b200 import initial_conditions as ic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos, vel, mass = ic.uniform_disk(n, seed=7, round6=False)
    sizes = [bh.shard_range(n, world, r) for r in range(world)]
    lo, hi = sizes[rank]
    # full-size arrays of which only the own slice is current (the others hold garbage, like on the device)
    rng = np.random.default_rng(100 + rank)
    p = rng.normal(size=(n, 2)); p[lo:hi] = pos[lo:hi]
    v = rng.normal(size=(n, 2)); v[lo:hi] = vel[lo:hi]

    def gather(a):                      # exchange_slices: one broadcast per owner
        for r, (l, h) in enumerate(sizes):
            t = torch.from_numpy(np.ascontiguousarray(a[l:h]))
            dist.broadcast(t, src=r)
            a[l:h] = t.numpy()
    gather(p); gather(v)
    keys = oracle.body_keys(p, oracle.root_bounds(p))        # the full build, identical on every rank
    order = np.argsort(keys, kind="stable")                  # sidx
    p, v, perm = p[order], v[order], order.copy()            # reorder_gather_kernel (perm: internal -> original)
    own_keys = keys[order][lo:hi]
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), perm=perm, kmin=own_keys.min(), kmax=own_keys.max(),
             pos_slice=p[lo:hi], vel_slice=v[lo:hi])
    # a getter: gather the (internal-order) slices, then un-permute
    p[:lo] = 0; p[hi:] = 0
    gather(p)
    out = np.empty_like(p); out[perm] = p
    np.save(os.path.join(out_dir, f"get_r{rank}.npy"), out)
    dist.destroy_process_group()


def test_gloo_world2_repartition_protocol(tmp_path):
    import torch.multiprocessing as mp
    from gpu_nbody_simulation_b200 import initial_conditions as ic
    n, world = 5001, 2
    mp.spawn(_repartition_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    pos, vel, mass = ic.uniform_disk(n, seed=7, round6=False)
    r = [np.load(os.path.join(tmp_path, f"r{k}.npz")) for k in range(world)]
    assert np.array_equal(r[0]["perm"], r[1]["perm"])                        # the same permutation on every rank
    assert np.array_equal(np.sort(r[0]["perm"]), np.arange(n))
    assert int(r[0]["kmax"]) <= int(r[1]["kmin"])                            # index slices = contiguous Morton ranges
    for k in range(world):
        assert np.array_equal(np.load(os.path.join(tmp_path, f"get_r{k}.npy")), pos)   # getters: the caller's order


@pytest.mark.gpu
def test_two_gpus_repartition_free_running_fp64():
    """20 free-running FP64 steps on 2 ranks with bodies in random order and one re-partition, against the single-GPU
    context (<= 1e-8) and the CPU oracle (<= 1e-8): tests/multi_gpu_check.py --repartition."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_check.py"), "--bodies", "100001",
           "--steps", "20", "--fp64", "--repartition"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "MULTI_GPU_CHECK PASS" in res.stdout and "re-partition(s)" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


# ------------------------------------------------------------------------------------------------
# The round-2 cell-sum exchange: every rank stores the sums of its NON-EMPTY finest cells into its slot of
# every rank's inbox; the consumer adds the slots in rank order and zeroes what it consumed.  gloo all_gather stands in
# for the peer stores; the result must equal the dense rank-ordered sum on every rank, bit for bit, step after step
# (the inbox is clean again without a memset).
# ------------------------------------------------------------------------------------------------
def _inbox_worker(rank, world, port, ncells, steps, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inbox = np.zeros((world, 4, ncells))                      # [source rank][count, m, mx, my][cell]
    results = []
    for s in range(steps):
        rng = np.random.default_rng(1000 * s + rank)
        cells = rng.choice(ncells, size=ncells // (world + 1), replace=False)     # this rank's occupied cells
        local = np.zeros((4, ncells))
        local[0, cells] = rng.integers(1, 50, size=len(cells))
        local[1:, cells] = rng.normal(size=(3, len(cells)))
        # push: only the non-empty cells travel (index + 4 values); every rank receives every rank's list
        idx = np.flatnonzero(local[0])
        payload = [None] * world
        dist.all_gather_object(payload, (idx, local[:, idx]))
        for src, (i, v) in enumerate(payload):
            inbox[src][:, i] = v
        # consume: rank order, skip empty slots, zero what was read
        total = np.zeros((4, ncells))
        for src in range(world):
            hit = inbox[src, 0] != 0.0
            total[:, hit] = total[:, hit] + inbox[src][:, hit]
            inbox[src][:, hit] = 0.0
        assert not inbox.any()                                 # clean for the next step
        dense = [None] * world
        dist.all_gather_object(dense, local)
        want = np.zeros((4, ncells))
        for src in range(world):
            want = want + dense[src]
        assert np.array_equal(total, want)
        results.append(total)
    np.save(os.path.join(out_dir, f"inbox_r{rank}.npy"), np.stack(results))
    dist.destroy_process_group()


def test_gloo_world2_sparse_inbox_exchange(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_inbox_worker, args=(world, _free_port(), 4096, 3, str(tmp_path)), nprocs=world, join=True)
    a, b = (np.load(os.path.join(tmp_path, f"inbox_r{k}.npy")) for k in range(world))
    assert np.array_equal(a, b)                                # identical reduced sums on every rank
