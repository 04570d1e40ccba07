"""CPU tests of the EXACT-LEAVES extension of the oracle (not reference behaviour):
multi-body cap-level leaves are applied as exact pair sums over their bodies, self excluded.  This is
the specification the CUDA flag BH_FLAG_EXACT_LEAVES is tested against (tests/test_gpu_exact_leaves.py).
"""
import numpy as np

import oracle
from conftest import golden_inputs


def per_body_rel(f, want):
    return np.linalg.norm(f - want, axis=1) / np.maximum(np.linalg.norm(want, axis=1), 1e-300)


def test_identical_to_reference_semantics_without_multi_body_leaves():
    rng = np.random.default_rng(3)
    pos = rng.uniform(-0.1, 0.1, size=(700, 2))
    mass = rng.uniform(0.1, 0.5, size=700)
    par = oracle.default_params(max_depth=17)            # deep enough: every cap-level leaf holds one body
    tree = oracle.Tree(pos, mass, par)
    nodes = tree.nodes()
    assert not np.any((nodes[:, 11] == -1) & (nodes[:, 6] > 0) & (nodes[:, 0] == -1)), "no multi-body leaves expected"
    f0, c0 = tree.forces()
    f1, c1 = tree.forces_exact_leaves()
    assert np.array_equal(f0, f1)
    assert c0 == c1


def test_exact_leaves_is_a_barnes_hut_approximation_of_the_direct_sum(shipped40k):
    """The reference's forces are dominated by the self-inclusive multi-body-leaf monopoles (SURVEY 0.10:
    rel-RMS 1.4e5 vs the direct sum); with exact leaves the same tree gives theta = 0.5 accuracy."""
    n = 12000
    pos, mass = shipped40k["pos"][:n], shipped40k["mass"][:n]
    tree = oracle.Tree(pos, mass)
    nt = oracle.max_threads()
    want = oracle.direct_forces(pos, mass, nthreads=nt)
    f_ref, c_ref = tree.forces(nthreads=nt)
    f_ex, c_ex = tree.forces_exact_leaves(nthreads=nt)
    rms = lambda f: float(np.sqrt(np.sum((f - want) ** 2) / np.sum(want ** 2)))
    assert rms(f_ref) > 1e3                                # the reference semantics, for the record
    assert rms(f_ex) < 1e-3                                # measured 2.3e-4 at 40 000 bodies
    per = per_body_rel(f_ex, want)
    assert np.median(per) < 2e-2 and np.percentile(per, 99) < 0.3
    # same walk, only the leaf handling differs: visits / opens equal, every body now skips itself once
    assert c_ex["visits"] == c_ref["visits"] and c_ex["opens"] == c_ref["opens"]
    assert c_ex["self_skips"] == n
    assert c_ex["interactions"] >= c_ref["interactions"]


def test_exact_leaves_small_cases_by_hand():
    # two bodies in the same cap-level cell (depth cap 1: the root is the only, multi-body, leaf) + a third
    pos = np.array([[0.0, 0.0], [3.0, 4.0], [-6.0, 8.0]])
    mass = np.array([2.0, 3.0, 5.0])
    par = oracle.default_params(max_depth=1, G=1.0, dist_eps=0.0)
    tree = oracle.Tree(pos, mass, par)
    assert tree.size == 1
    f, cnt = tree.forces_exact_leaves()
    want = np.zeros((3, 2))
    for i in range(3):
        for j in range(3):
            if i != j:
                d = pos[j] - pos[i]
                r = np.hypot(*d)
                want[i] += mass[i] * mass[j] * d / r ** 3
    assert np.allclose(f, want, rtol=1e-14, atol=0)
    assert cnt["interactions"] == 6 and cnt["self_skips"] == 3
    # coincident bodies in one leaf: d2 == 0 -> inf * 0 = NaN, the reference's own arithmetic (A.6)
    pos2, vel2, mass2, _ = golden_inputs("tiny_2_coincident")
    f2, _ = oracle.Tree(pos2, mass2).forces_exact_leaves()
    assert np.isnan(f2).all()
