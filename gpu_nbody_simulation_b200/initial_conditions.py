"""Seeded synthetic initial conditions (host side, numpy).

Replaces the reference's non-reproducible random init (``initializeCpu`` / ``initializeGpu``,
project.cu:298-341, seeded with ``time``) with counter-based, seeded generators, and reads /
writes the reference's three text files (``loadSimulationDataFromText`` project.cu:103-161;
writers project.cu:230-282).  Value ranges follow project.cu:30-35.

``round6=True`` passes every value through the reference writers' formatting (default ostream
precision = 6 significant digits, i.e. ``%.6g``) so that the in-memory FP64 arrays and the text
files describe exactly the same doubles.
"""
from __future__ import annotations

import os

import numpy as np

LOWER_M, HIGHER_M = 1e-1, 5e-1      # project.cu:30-31
LOWER_P, HIGHER_P = -1e-1, 1e-1     # project.cu:32-33
LOWER_V, HIGHER_V = -1e-4, 1e-4     # project.cu:34-35


def _round6(a: np.ndarray) -> np.ndarray:
    flat = np.ascontiguousarray(a, dtype=np.float64).ravel()
    txt = np.char.mod("%.6g", flat)
    return txt.astype(np.float64).reshape(a.shape)


def _finish(pos, vel, mass, round6):
    if round6:
        pos, vel, mass = _round6(pos), _round6(vel), _round6(mass)
    return (np.ascontiguousarray(pos, dtype=np.float64), np.ascontiguousarray(vel, dtype=np.float64),
            np.ascontiguousarray(mass, dtype=np.float64))


def _masses_log_uniform(rng, n, lo, hi):
    # project.cu:86-89 / :99-101: 10 ** (log10(lo) + u * (log10(hi) - log10(lo)))
    return np.power(10.0, np.log10(lo) + rng.random(n) * (np.log10(hi) - np.log10(lo)))


def uniform_square(n: int, seed: int = 12345, round6: bool = True):
    """The reference's own distribution: positions U(+-0.1)^2 (project.cu:32-33, :93-95)."""
    rng = np.random.Generator(np.random.Philox(seed))
    pos = LOWER_P + rng.random((n, 2)) * (HIGHER_P - LOWER_P)
    vel = LOWER_V + rng.random((n, 2)) * (HIGHER_V - LOWER_V)
    mass = _masses_log_uniform(rng, n, LOWER_M, HIGHER_M)
    return _finish(pos, vel, mass, round6)


def uniform_disk(n: int, seed: int = 12345, radius: float = 0.1, round6: bool = True):
    """BASELINE.json config 2/4: uniform disk r = R sqrt(u1), phi = 2 pi u2."""
    rng = np.random.Generator(np.random.Philox(seed))
    r = radius * np.sqrt(rng.random(n))
    phi = 2.0 * np.pi * rng.random(n)
    pos = np.stack([r * np.cos(phi), r * np.sin(phi)], axis=1)
    vel = LOWER_V + rng.random((n, 2)) * (HIGHER_V - LOWER_V)
    mass = _masses_log_uniform(rng, n, LOWER_M, HIGHER_M)
    return _finish(pos, vel, mass, round6)


def plummer_2d(n: int, seed: int = 12345, a: float = 0.02, rmax: float = 0.1, round6: bool = True):
    """BASELINE.json config 3: Plummer sphere r = a / sqrt(u^(-2/3) - 1), isotropic direction,
    z dropped, truncated at 3-D radius <= rmax (rejection by resampling)."""
    rng = np.random.Generator(np.random.Philox(seed))
    out = np.empty((n, 2))
    got = 0
    while got < n:
        m = n - got
        u = rng.random(m)
        u = np.where(u <= 0.0, 0.5, u)
        r = a / np.sqrt(np.power(u, -2.0 / 3.0) - 1.0)
        cz = 2.0 * rng.random(m) - 1.0
        ph = 2.0 * np.pi * rng.random(m)
        s = np.sqrt(1.0 - cz * cz)
        keep = r <= rmax
        k = int(keep.sum())
        out[got:got + k, 0] = (r * s * np.cos(ph))[keep]
        out[got:got + k, 1] = (r * s * np.sin(ph))[keep]
        got += k
    vel = LOWER_V + rng.random((n, 2)) * (HIGHER_V - LOWER_V)
    mass = _masses_log_uniform(rng, n, LOWER_M, HIGHER_M)
    return _finish(out, vel, mass, round6)


GENERATORS = {"uniform_square": uniform_square, "uniform_disk": uniform_disk, "plummer_2d": plummer_2d}


# ---------------------------------------------------------------------------------------------
# The reference's three text files (cwd-relative names are the reference's: project.cu:1065).
# ---------------------------------------------------------------------------------------------
def write_init_files(directory: str, pos, vel, mass):
    """masses_init.txt: one value per line; positions/velocities: 'x y' per line, %.6g
    (project.cu:242, :276)."""
    os.makedirs(directory, exist_ok=True)
    np.savetxt(os.path.join(directory, "masses_init.txt"), np.asarray(mass), fmt="%.6g")
    np.savetxt(os.path.join(directory, "positions_init.txt"), np.asarray(pos).reshape(-1, 2), fmt="%.6g")
    np.savetxt(os.path.join(directory, "velocities_init.txt"), np.asarray(vel).reshape(-1, 2), fmt="%.6g")


def read_init_files(directory: str, n_bodies: int):
    """First ``n_bodies`` lines of each file; too few lines raises (project.cu:121-124, :137-140)."""
    def load(name, cols):
        a = np.loadtxt(os.path.join(directory, name), max_rows=n_bodies, ndmin=2)
        if a.shape[0] < n_bodies:
            raise RuntimeError(f"Not enough entries in file: {name}")
        return np.ascontiguousarray(a[:, :cols], dtype=np.float64)
    mass = load("masses_init.txt", 1).reshape(-1)
    pos = load("positions_init.txt", 2)
    vel = load("velocities_init.txt", 2)
    return pos, vel, mass
