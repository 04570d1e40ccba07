"""Resident Morton order (single rank): the engine re-sorts its body arrays physically from time to time; the C-ABI
keeps the caller's ORIGINAL body order on every setter and getter (include/bh.h), and results equal the un-reordered
engine up to the summation order inside shared cap-level cells."""
import os

import numpy as np
import pytest

import oracle
from gpu_nbody_simulation_b200 import Simulation
from gpu_nbody_simulation_b200 import initial_conditions as ic

pytestmark = pytest.mark.gpu
GENTLE = dict(G=6.67e-11 * 1e-6)      # dynamics that stay comparable over many steps


def rel_rms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def test_many_steps_with_and_without_reordering_agree():
    n = 150_000
    pos, vel, mass = ic.uniform_disk(n, seed=3, round6=False)
    out = {}
    for flag in ("1", "0"):
        os.environ["BH_REORDER"] = flag
        try:
            with Simulation(n, **GENTLE) as sim:
                sim.set_bodies(pos, vel, mass)
                sim.step(40)                      # re-sorted before step 2, then every 16 steps
                assert sim.counters()["reorders"] == (3 if flag == "1" else 0)
                out[flag] = (sim.positions(), sim.velocities(), sim.forces(), sim.accelerations())
        finally:
            del os.environ["BH_REORDER"]
    for a, b in zip(out["1"], out["0"]):
        assert rel_rms(a, b) <= 1e-6
    assert rel_rms(out["1"][0] - pos, out["0"][0] - pos) <= 1e-5     # the displacement itself, not only the position


def test_getters_and_setters_keep_the_original_order_after_a_reorder():
    n = 60_000
    pos, vel, mass = ic.uniform_disk(n, seed=5, round6=False)
    with Simulation(n, **GENTLE) as sim:
        sim.set_bodies(pos, vel, mass)
        sim.step(3)                               # arrays are in resident order now
        p, v = sim.positions(), sim.velocities()
        assert rel_rms(p, pos) < 1e-1 and not np.array_equal(p, pos)          # same bodies, slightly moved, same order
        sim.build_tree()
        tree = oracle.Tree(p, mass, oracle.default_params(G=GENTLE["G"]))
        assert np.array_equal(sim.bounds(), oracle.root_bounds(p))
        assert np.array_equal(sim.body_keys(), oracle.body_keys(p, sim.bounds()))
        order = sim.sorted_order().astype(np.int64)
        assert np.array_equal(np.sort(order), np.arange(n))
        keys = sim.body_keys()[order]
        assert (np.diff(keys.astype(np.int64)) >= 0).all()
        got, want = sim.tree(), tree.canonical()
        assert got.shape == want.shape
        assert np.array_equal(got[:, [0, 1, 2, 3, 4, 8, 9]], want[:, [0, 1, 2, 3, 4, 8, 9]])   # topology, boxes, occupants
        assert np.allclose(got[:, 5:8], want[:, 5:8], rtol=1e-12, atol=1e-300)                 # in-cell summation order
        sim.compute_forces()
        f_ref, _ = tree.forces(nthreads=oracle.max_threads())
        assert rel_rms(sim.forces(), f_ref) <= 1e-5
        # setters in the caller's order
        newp = p[::-1].copy()
        sim.set_positions(newp)
        assert np.array_equal(sim.positions(), newp)
        sim.set_velocities(2.0 * v)
        assert np.array_equal(sim.velocities(), 2.0 * v)


def test_snapshot_restore_and_restart_steps_under_reordering():
    n = 100_000
    pos, vel, mass = ic.uniform_disk(n, seed=9, round6=False)
    with Simulation(n, **GENTLE) as sim, Simulation(n, **GENTLE) as ref:
        sim.set_bodies(pos, vel, mass)
        sim.snapshot()                            # taken in resident order
        assert np.array_equal(sim.positions(), pos) and np.array_equal(sim.velocities(), vel)
        sim.step_from_snapshot(3)
        ref.set_bodies(pos, vel, mass)
        ref.step(1)
        assert rel_rms(sim.positions() - pos, ref.positions() - pos) <= 1e-5
        sim.step(20)                              # re-sorts the arrays (and the snapshot with them)
        sim.restore()
        assert np.array_equal(sim.positions(), pos) and np.array_equal(sim.velocities(), vel)
