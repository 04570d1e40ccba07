"""Seeded initial conditions (csrc/generate.cu): the host twin of the device generator,
the raw Philox4x32-10 against the Random123 known-answer vectors, and the writers of the reference's three
text files.  CPU only; the device kernel is compared with the host twin in tests/test_gpu_generate.py."""
import os

import numpy as np
import pytest

import gpu_nbody_simulation_b200 as bh
from gpu_nbody_simulation_b200 import initial_conditions as ic


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors, "philox4x32 10"
    kat = [([0x00000000] * 4, [0x00000000] * 2, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert bh.philox4x32_10(ctr, key) == want


@pytest.mark.parametrize("kind", ["uniform_square", "uniform_disk", "plummer_2d"])
def test_value_ranges_and_reproducibility(kind):
    n = 50000
    pos, vel, mass = bh.generate_host(kind, n, seed=12345)
    assert np.isfinite(pos).all() and np.isfinite(vel).all() and np.isfinite(mass).all()
    assert (mass >= 0.1).all() and (mass < 0.5).all()                      # project.cu:30-31
    assert (np.abs(vel) <= 1e-4).all()                                     # project.cu:34-35
    assert abs(np.log(mass).mean() - np.log(np.sqrt(0.05))) < 0.01         # log-uniform
    r = np.hypot(pos[:, 0], pos[:, 1])
    if kind == "uniform_square":
        assert (np.abs(pos) <= 0.1).all() and abs(pos.mean()) < 1e-3       # project.cu:32-33
        assert abs(np.mean(pos ** 2) - 0.1 ** 2 / 3) < 1e-4
    elif kind == "uniform_disk":
        assert (r <= 0.1).all()
        assert abs(np.mean((r / 0.1) ** 2) - 0.5) < 5e-3                   # r^2 uniform on [0, R^2]
        assert abs(np.mean(pos[:, 0] * pos[:, 1])) < 5e-5
    else:
        assert (r <= 0.1).all()
        want = ic.plummer_2d(n, seed=1, round6=False)[0]                   # same distribution, numpy generator
        rw = np.hypot(want[:, 0], want[:, 1])
        for q in (25, 50, 75, 95):
            assert abs(np.percentile(r, q) / np.percentile(rw, q) - 1) < 0.04
    # pure function of (seed, body index): any slice reproduces, another seed differs
    part = bh.generate_host(kind, 1000, seed=12345, first=n - 1000)
    assert all(np.array_equal(a, b[n - 1000:]) for a, b in zip(part, (pos, vel, mass)))
    other = bh.generate_host(kind, 1000, seed=12346)
    assert not np.array_equal(other[0], pos[:1000])


def test_init_files_match_the_python_writer_and_round_trip(tmp_path):
    n = 3000
    pos, vel, mass = bh.generate_host("uniform_disk", n, seed=5)
    a, b = str(tmp_path / "c"), str(tmp_path / "py")
    bh.write_init_files(a, pos, vel, mass)                                 # C writer ("%g", project.cu:236-281)
    ic.write_init_files(b, pos, vel, mass)                                 # numpy "%.6g"
    for name in ("masses_init.txt", "positions_init.txt", "velocities_init.txt"):
        assert open(os.path.join(a, name), "rb").read() == open(os.path.join(b, name), "rb").read(), name
    p2, v2, m2 = bh.load_text(a, n)                                        # loadSimulationDataFromText
    pr, vr, mr = bh.generate_host("uniform_disk", n, seed=5, round6=True)
    assert np.array_equal(p2, pr) and np.array_equal(v2, vr) and np.array_equal(m2, mr)
    import ctypes as C
    dp = C.POINTER(C.c_double)
    bad = str(tmp_path / "missing_dir" / "m.txt").encode()
    rc = bh.lib().bh_write_init_files(bad, bad, bad, n, mass.ctypes.data_as(dp), pos.ctypes.data_as(dp), vel.ctypes.data_as(dp))
    assert rc == -4 and b"Failed to open file for writing masses." in bh.lib().bh_last_error()   # project.cu:239
