"""Adversarial inputs through the CUDA path (tree bit-exact, FP64-mode forces, NaN pattern) — the same cases
tests/test_oracle.py pins against the live reference on the CPU.

NOT YET RUN ON A GPU (written after round 1's GPU budget was spent); skipped unless BH_TEST_UNVALIDATED=1.
"""
import os

import numpy as np
import pytest

import oracle
from gpu_nbody_simulation_b200 import Simulation
from test_oracle import _edge_case

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("BH_TEST_UNVALIDATED") != "1",
                                 reason="edge-case GPU tests not yet validated on a GPU (set BH_TEST_UNVALIDATED=1)")]


@pytest.mark.parametrize("kind", ["collinear", "all_coincident", "two_far_clusters", "tiny_separations", "lattice",
                                  "zero_and_tiny_masses"])
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
