"""Adversarial inputs through the CUDA path (tree bit-exact, FP64-mode forces, NaN pattern) — the same cases
tests/test_oracle.py pins against the live reference on the CPU.
"""
import numpy as np
import pytest

import oracle
from gpu_nbody_simulation_b200 import Simulation
from test_oracle import _edge_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["collinear", "all_coincident", "two_far_clusters", "tiny_separations", "lattice",
                                  "tiny_masses"])
def test_edge_case_tree_and_fp64_forces(kind):
    n = 1000
    pos, mass = _edge_case(kind, n, np.random.default_rng(len(kind)))
    vel = np.zeros((n, 2))
    tree = oracle.Tree(pos, mass)
    want, cnt = tree.forces()
    with Simulation(n, fp64=True, counters=True, exact_leaf_max=1 << 20) as sim:
        sim.set_bodies(pos, vel, mass)
        sim.build_tree()
        assert np.array_equal(sim.bounds(), oracle.root_bounds(pos))
        assert sim.tree_size() == tree.size
        assert np.array_equal(sim.tree(), tree.canonical(), equal_nan=True)
        sim.compute_forces()
        f = sim.forces()
        assert np.array_equal(np.isnan(f), np.isnan(want)), "NaN pattern (body exactly at a leaf COM, project.cu:765-772)"
        ok = np.isfinite(want).all(axis=1)
        if ok.any() and np.abs(want[ok]).max() > 0:
            err = np.sqrt(((f[ok] - want[ok]) ** 2).sum() / (want[ok] ** 2).sum())
            assert err <= 1e-12, err
        c = sim.counters()
        assert c["interactions"] == cnt["interactions"] and c["visits"] == cnt["visits"] and c["opens"] == cnt["opens"]


def test_exactly_zero_masses_are_detected_not_reproduced():
    """Bodies of mass exactly 0 are "ghosts" in the reference: a leaf holding one counts as empty and is overwritten by
    the next arrival (is_empty_leaf tests TOTAL_MASS == 0, project.cu:393-405), so its topology depends on the insertion
    order.  The engine builds the order-independent tree (every body counts) and REPORTS the condition instead:
    bh_counters.reserved[0] = number of zero-mass bodies handed over.  The reference's own generators
    draw masses from [0.1, 0.5] (project.cu:30-31) and never produce one."""
    n = 1000
    pos, mass = _edge_case("zero_and_tiny_masses", n, np.random.default_rng(20))
    with Simulation(n, fp64=True, counters=True) as sim:
        sim.set_bodies(pos, np.zeros((n, 2)), mass)
        sim.build_tree()
        assert sim.counters()["zero_mass_bodies"] == int((mass == 0.0).sum()) > 0
        assert np.array_equal(sim.bounds(), oracle.root_bounds(pos))
        # the tree over the bodies as handed over is the reference's tree when the zero masses are made non-zero
        mass2 = np.where(mass == 0.0, 1e-300, mass)
        tree = oracle.Tree(pos, mass2)
        assert sim.tree_size() == tree.size
