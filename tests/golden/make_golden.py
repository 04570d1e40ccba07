"""Generates tests/golden/*.npz by running the REFERENCE's own CPU functions (unmodified
implementation/project.cu, compiled where it lies by oracle/build_ref.sh, driven by
oracle/ref_harness.cu).  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array in the fixtures is raw FP64 output of buildTree / computeForces / update* of the
reference; large cases store SHA-256 digests of the raw bytes plus strided subsamples.
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from gpu_nbody_simulation_b200 import initial_conditions as ic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF_DIR = "/root/reference/implementation"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ensure(n):
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh"), str(n)])


def full_case(name, pos, vel, mass, steps, store_inputs=True):
    n = mass.shape[0]
    ensure(n)
    recs, tim = oracle.run_ref(pos, vel, mass, steps=steps)
    d = {"n": n, "steps": steps}
    if store_inputs:
        d.update(pos=pos, vel=vel, mass=mass)
    for s in range(steps):
        d[f"tree{s}"] = recs[("tree", s)].reshape(-1, 12)
        for k in ("forces", "acc", "vel", "pos"):
            d[f"{k}{s}" if k != "vel" and k != "pos" else f"{k}_after{s}"] = recs[(k, s)].reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "nodes per step", [t["nodes"] for t in tim])


def digest_case(name, pos, vel, mass, steps, sub, store_inputs):
    n = mass.shape[0]
    ensure(n)
    recs, tim = oracle.run_ref(pos, vel, mass, steps=steps)
    d = {"n": n, "steps": steps, "sub": sub, "inputs_sha": sha(np.concatenate([mass, pos.ravel(), vel.ravel()]))}
    if store_inputs:
        d.update(pos=pos, vel=vel, mass=mass)
    for s in range(steps):
        tree = recs[("tree", s)].reshape(-1, 12)
        d[f"nodes{s}"] = tree.shape[0]
        d[f"bounds{s}"] = tree[0, 7:11]
        d[f"root_mass_com{s}"] = tree[0, 4:7]
        d[f"tree_sha{s}"] = sha(tree)
        for k in ("forces", "acc", "vel", "pos"):
            a = recs[(k, s)].reshape(-1, 2)
            d[f"{k}_sha{s}"] = sha(a)
            d[f"{k}_sub{s}"] = a[::sub].copy()
        f = recs[("forces", s)].reshape(-1, 2)
        d[f"forces_sumsq{s}"] = float(np.sum(f * f))
    d["ref_timing_us"] = np.array([[t["build_us"], t["force_us"], t["update_us"]] for t in tim])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "nodes per step", [t["nodes"] for t in tim], "timing", tim[0])


def main():
    # 1. The reference's shipped initial conditions, first 40 000 bodies (BASELINE config 1).
    N1 = 40000
    pos = np.loadtxt(os.path.join(REF_DIR, "positions_init.txt"))[:N1]
    vel = np.loadtxt(os.path.join(REF_DIR, "velocities_init.txt"))[:N1]
    mass = np.loadtxt(os.path.join(REF_DIR, "masses_init.txt"))[:N1]
    digest_case("shipped_40000", pos, vel, mass, steps=3, sub=16, store_inputs=True)
    # 2. A prefix small enough to store every intermediate in full (inputs = prefix of case 1).
    full_case("shipped_2048", pos[:2048].copy(), vel[:2048].copy(), mass[:2048].copy(), steps=3, store_inputs=False)
    # 3. Clustered synthetic input with many bodies sharing finest cells and exact duplicates.
    rng = np.random.Generator(np.random.Philox(7))
    c = rng.normal(0.0, 2e-4, size=(1000, 2))
    c[100:200] = c[100]                       # 100 coincident bodies
    c[900:] = rng.uniform(-0.1, 0.1, size=(100, 2))
    cv = rng.uniform(-1e-4, 1e-4, size=(1000, 2))
    cm = np.power(10.0, rng.uniform(-1, np.log10(0.5), size=1000))
    full_case("clustered_1000", c, cv, cm, steps=2)
    # 4. Degenerate sizes: a single body, two coincident bodies (maxDim == 0 -> 1e-6 padding), five.
    full_case("tiny_1", np.array([[0.25, -0.5]]), np.array([[1e-5, 2e-5]]), np.array([0.3]), steps=2)
    full_case("tiny_2_coincident", np.array([[0.01, 0.02], [0.01, 0.02]]), np.zeros((2, 2)), np.array([0.2, 0.4]), steps=2)
    p5 = np.array([[0.0, 0.0], [1e-3, 0.0], [1e-3, 0.0], [-0.05, 0.08], [0.09, -0.02]])
    full_case("tiny_5", p5, np.zeros((5, 2)), np.array([0.1, 0.2, 0.3, 0.4, 0.5]), steps=2)
    # 5. BASELINE config 2: one million bodies, uniform disk (inputs regenerated from the seed).
    pos, vel, mass = ic.uniform_disk(1_000_000, seed=12345)
    digest_case("disk_1000000", pos, vel, mass, steps=2, sub=1000, store_inputs=False)


if __name__ == "__main__":
    main()
