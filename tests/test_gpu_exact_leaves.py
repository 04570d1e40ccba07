"""GPU tests of BH_FLAG_EXACT_LEAVES (extension) against the oracle's
bho_compute_forces_exact_leaves.
"""
import numpy as np
import pytest

import oracle
from conftest import golden_inputs
from gpu_nbody_simulation_b200 import BhError, Simulation

pytestmark = pytest.mark.gpu


def rel_rms(a, b):
    ok = np.isfinite(b).all(axis=-1)
    return float(np.sqrt(np.sum((a[ok] - b[ok]) ** 2) / max(np.sum(b[ok] ** 2), 1e-300)))


@pytest.mark.parametrize("name", ["shipped_2048", "clustered_1000", "tiny_5", "tiny_1", "shipped_40000"])
def test_forces_match_the_oracle_extension(name):
    pos, vel, mass, _ = golden_inputs(name)
    tree = oracle.Tree(pos, mass)
    want, cnt = tree.forces_exact_leaves(nthreads=oracle.max_threads())
    with Simulation(len(mass), fp64=True, counters=True, exact_leaves=True, exact_leaf_max=1 << 20) as sim:
        sim.set_bodies(pos, vel, mass)
        sim.build_tree()
        sim.compute_forces()
        f = sim.forces()
        assert np.array_equal(np.isnan(f), np.isnan(want))
        assert rel_rms(f, want) <= 1e-12
        c = sim.counters()
        assert c["interactions"] == cnt["interactions"] and c["visits"] == cnt["visits"] and c["opens"] == cnt["opens"]
    with Simulation(len(mass), counters=True, exact_leaves=True) as sim:
        sim.set_bodies(pos, vel, mass)
        sim.build_tree()
        sim.compute_forces()
        f = sim.forces()
        ok = np.isfinite(want).all(axis=1)
        assert rel_rms(f[ok], want[ok]) <= 1e-5        # FP32 traversal, same bar as the reference-semantics mode
        per = np.linalg.norm(f[ok] - want[ok], axis=1) / np.maximum(np.linalg.norm(want[ok], axis=1), 1e-300)
        assert np.median(per) <= 1e-5
    # the pair kernel's exact-leaves path (2: two bodies per lane, packed) and the list kernel's (9: members staged in
    # the warp-local frame)
    for bpl in (2, 9):
        for exact_eps in (False, True):
            with Simulation(len(mass), exact_leaves=True, bodies_per_lane=bpl, exact_eps=exact_eps) as sim:
                sim.set_bodies(pos, vel, mass)
                sim.build_tree()
                sim.compute_forces()
                f = sim.forces()
                ok = np.isfinite(want).all(axis=1)
                assert rel_rms(f[ok], want[ok]) <= 1e-5, (bpl, exact_eps)
                per = np.linalg.norm(f[ok] - want[ok], axis=1) / np.maximum(np.linalg.norm(want[ok], axis=1), 1e-300)
                assert np.median(per) <= 1e-5, (bpl, exact_eps)


def test_against_the_direct_sum(shipped40k):
    n = 12000
    pos, vel, mass = shipped40k["pos"][:n], shipped40k["vel"][:n], shipped40k["mass"][:n]
    want = oracle.direct_forces(pos, mass, nthreads=oracle.max_threads())
    with Simulation(n, exact_leaves=True) as sim:
        sim.set_bodies(pos, vel, mass)
        sim.build_tree()
        sim.compute_forces()
        assert rel_rms(sim.forces(), want) < 1e-3     # oracle extension: 2.3e-4 at 40 000 bodies (theta = 0.5)


def test_root_is_the_multi_body_leaf_and_whole_step():
    rng = np.random.default_rng(8)
    n = 300
    pos = rng.uniform(-1, 1, size=(n, 2)); vel = rng.uniform(-1e-4, 1e-4, size=(n, 2)); mass = rng.uniform(0.1, 0.5, size=n)
    par = oracle.default_params(max_depth=1)
    want, _ = oracle.Tree(pos, mass, par).forces_exact_leaves()
    for fp64, tol, bpl in ((True, 1e-12, 0), (False, 1e-5, 0), (False, 1e-5, 9)):
        with Simulation(n, fp64=fp64, exact_leaves=True, max_depth=1, exact_leaf_max=1 << 20, bodies_per_lane=bpl) as sim:
            sim.set_bodies(pos, vel, mass)
            sim.step(1)                                # fused integrator epilogue of the exact-leaves kernels
            assert rel_rms(sim.forces(), want) <= tol
            acc, v, p = oracle.update(want, mass, vel, pos, par.dt)
            assert rel_rms(sim.positions(), p) <= tol and rel_rms(sim.velocities(), v) <= tol


def test_multi_rank_exact_leaves_needs_the_communicator():
    """On a multi-rank context the flag makes every step all-gather the positions (a leaf's bodies may live on other
    ranks): without an attached NCCL communicator the step fails loudly (the 2-GPU run is tests/multi_gpu_check.py
    --exact-leaves)."""
    rng = np.random.default_rng(3)
    n = 1000
    with Simulation(n, exact_leaves=True, rank=0, n_ranks=2) as sim:
        sim.set_bodies(rng.uniform(-1, 1, (n, 2)), np.zeros((n, 2)), rng.uniform(0.1, 0.5, n))
        with pytest.raises(BhError):
            sim.step(1)
    # the refusal must not have loaded the system's libnccl: a later `import torch` (its own bundled libnccl, same SONAME)
    # would fail — which is how this was found (the GPU suite imports torch in later tests)
    import torch  # noqa: F401
    loaded = [ln.split()[-1] for ln in open("/proc/self/maps") if "libnccl" in ln]
    assert all("nvidia" in p or "torch" in p for p in loaded), loaded


@pytest.mark.parametrize("bodies_per_lane", [0, 2, 9])
def test_whole_step_many_blocks_reads_pre_step_positions(bodies_per_lane):
    """The member loop reads OTHER bodies' positions; an in-place fused integrator in a block that finished earlier
    would hand it post-step positions (a race that a few co-resident blocks never show).  bh_step therefore runs the
    forces-only kernel followed by one integrator launch when the flag is set: 200 000 bodies = 1 500+ blocks."""
    from gpu_nbody_simulation_b200 import initial_conditions as ic
    n = 200_000
    pos, vel, mass = ic.uniform_disk(n, seed=7, round6=False)
    want, _ = oracle.Tree(pos, mass).forces_exact_leaves(nthreads=oracle.max_threads())
    acc_w, vel_w, pos_w = oracle.update(want, mass, vel, pos, 1.0)
    with Simulation(n, exact_leaves=True, bodies_per_lane=bodies_per_lane) as sim:
        sim.set_bodies(pos, vel, mass)
        sim.step(1)
        ok = np.isfinite(want).all(axis=1)
        assert rel_rms(sim.forces()[ok], want[ok]) <= 1e-5
        assert rel_rms(sim.accelerations()[ok], acc_w[ok]) <= 1e-5
        assert rel_rms(sim.positions()[ok] - pos[ok], pos_w[ok] - pos[ok]) <= 1e-5
        f1 = sim.forces().copy()
        sim.set_bodies(pos, vel, mass)
        sim.step(1)                                  # deterministic: a race would differ from run to run
        assert np.array_equal(sim.forces(), f1)
