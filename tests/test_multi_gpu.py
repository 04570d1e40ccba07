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


@pytest.mark.gpu
def test_two_gpus_match_single_gpu_and_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "--bodies", "100001",
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
    port = 29600 + (n % 50)
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
