"""Device-side seeded generator (bh_generate, csrc/generate.cu) against its host twin.

The host twin, the Philox known answers and the file writers are tested on the CPU in tests/test_generate.py.
"""
import numpy as np
import pytest

import gpu_nbody_simulation_b200 as bh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["uniform_square", "uniform_disk", "plummer_2d"])
def test_device_generator_matches_host_twin(kind):
    n = 100003
    pos, vel, mass = bh.generate_host(kind, n, seed=99)
    with bh.Simulation(n) as sim:
        sim.generate(kind, seed=99)
        p, v = sim.positions(), sim.velocities()
        assert np.array_equal(v, vel)                          # multiply + add only, uncontracted on both sides
        if kind == "uniform_square":
            assert np.array_equal(p, pos)
        else:                                                  # sqrt / sin / cos / pow: last-bit differences
            assert np.allclose(p, pos, rtol=1e-13, atol=1e-17)
        sim.step(1)                                            # masses are in place: a step runs and stays finite
        sim.build_tree()
        assert sim.tree_size() > 1
        f = sim.forces()
        assert np.isfinite(f).sum() > 0.99 * f.size
