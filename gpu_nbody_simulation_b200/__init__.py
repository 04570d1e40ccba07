"""B200-native 2-D Barnes-Hut engine: drop-in for the simulation path of
DavidSevic/gpu-nbody-simulation (``implementation/project.cu``).

Python is only the host-side mirror used by the tests and ``bench.py``; the product is the
C-ABI shared library ``libbh.so`` (``include/bh.h``) built from ``csrc/*.cu`` for sm_100a.
There is NO CPU fallback: loading fails loudly when the library has not been built, and
creating a :class:`Simulation` fails when no CUDA device is present.

    sim = Simulation(n_bodies=40000)              # defaults = the reference's constants
    sim.set_bodies(pos, vel, mass)                # project.cu:943-945
    sim.step(10)                                  # loop body project.cu:955-1011, N_SIMULATIONS times
    pos = sim.positions()                         # project.cu:1010

See DESIGN.md for the architecture and INTEGRATION.md for the reference-side binding.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import initial_conditions  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BH_LIB", os.path.join(_HERE, "libbh.so"))   # BH_LIB: A/B builds of the same ABI

BH_FLAG_FP64_TRAVERSAL = 1 << 0
BH_FLAG_COUNTERS = 1 << 1
BH_FLAG_NO_GRAPH = 1 << 2
BH_FLAG_EXACT_EPS = 1 << 3
BH_FLAG_EXACT_LEAVES = 1 << 4   # extension (not reference behaviour), see include/bh.h

# every symbol include/bh.h declares (checked by tests/test_abi.py against the header text)
ABI_SYMBOLS = (
    "bh_last_error", "bh_abi_version", "bh_default_params", "bh_create", "bh_destroy", "bh_nccl_unique_id",
    "bh_attach_nccl", "bh_comm_handle", "bh_attach_peers", "bh_shard_range", "bh_set_bodies", "bh_set_positions", "bh_set_velocities", "bh_snapshot",
    "bh_restore", "bh_step", "bh_step_host", "bh_step_from_snapshot", "bh_build_tree", "bh_compute_forces", "bh_integrate",
    "bh_synchronize", "bh_get_positions", "bh_get_velocities", "bh_get_accelerations", "bh_get_forces",
    "bh_get_bounds", "bh_get_body_keys", "bh_get_sorted_order", "bh_get_tree_size", "bh_get_tree",
    "bh_dump_quadtree", "bh_get_counters", "bh_set_profiling", "bh_get_timers", "bh_reset_timers",
    "bh_last_step_ms", "bh_direct_forces", "bh_load_text", "bh_append_positions_txt", "bh_measure_fp32_peak",
    "bh_generate", "bh_generate_host", "bh_philox4x32_10", "bh_write_init_files",
    "bh_trajectory_begin", "bh_trajectory_record", "bh_trajectory_end",
)
GENERATOR_KINDS = {"uniform_square": 0, "uniform_disk": 1, "plummer_2d": 2}   # BH_GEN_* of include/bh.h


class BhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbh error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """struct bh_params (include/bh.h)."""

    _fields_ = [
        ("n_bodies", C.c_int64), ("G", C.c_double), ("dt", C.c_double), ("theta", C.c_double),
        ("dist_eps", C.c_double), ("mass_eps", C.c_double), ("pad_frac", C.c_double), ("pad_fallback", C.c_double),
        ("max_depth", C.c_int32), ("device", C.c_int32), ("flags", C.c_uint32), ("exact_leaf_max", C.c_int32),
        ("rank", C.c_int32), ("n_ranks", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class Counters(C.Structure):
    _fields_ = [("interactions", C.c_uint64), ("visits", C.c_uint64), ("opens", C.c_uint64),
                ("warp_steps", C.c_uint64), ("nodes", C.c_uint64), ("heavy_cells", C.c_uint64),
                ("zero_mass_bodies", C.c_uint64), ("reorders", C.c_uint64)]


class Timers(C.Structure):
    _fields_ = [("bounds_keys_us", C.c_double), ("sort_us", C.c_double), ("build_us", C.c_double),
                ("traverse_us", C.c_double), ("integrate_us", C.c_double), ("exchange_us", C.c_double),
                ("total_us", C.c_double), ("steps", C.c_uint64), ("kernel_launches", C.c_uint64)]


def build(verbose: bool = False) -> str:
    """Compile libbh.so in-tree with nvcc for sm_100a (see Makefile)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", _HERE, "-j8"], stdout=out)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build the CUDA library first "
                          "(python -c 'import __graft_entry__ as g; g.build()' or make -C gpu_nbody_simulation_b200). "
                          "This package has no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    dp, u32p, i64p, vp = C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_int64), C.c_void_p
    L.bh_last_error.restype = C.c_char_p
    L.bh_abi_version.restype = C.c_int
    L.bh_default_params.argtypes = [C.POINTER(Params)]
    L.bh_default_params.restype = None
    L.bh_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.bh_destroy.argtypes = [vp]
    L.bh_nccl_unique_id.argtypes = [vp]
    L.bh_attach_nccl.argtypes = [vp, vp]
    L.bh_comm_handle.argtypes = [vp, vp]
    L.bh_attach_peers.argtypes = [vp, vp, C.c_int32]
    L.bh_shard_range.argtypes = [C.c_int64, C.c_int32, C.c_int32, i64p, i64p]
    L.bh_set_bodies.argtypes = [vp, vp, vp, vp]
    L.bh_set_positions.argtypes = [vp, vp]
    L.bh_set_velocities.argtypes = [vp, vp]
    for name in ("bh_snapshot", "bh_restore", "bh_build_tree", "bh_compute_forces", "bh_integrate",
                 "bh_synchronize", "bh_reset_timers"):
        getattr(L, name).argtypes = [vp]
    L.bh_step.argtypes = [vp, C.c_int32]
    L.bh_step_from_snapshot.argtypes = [vp, C.c_int32]
    L.bh_step_host.argtypes = [vp, vp, vp, vp, vp]
    for name in ("bh_get_positions", "bh_get_velocities", "bh_get_accelerations", "bh_get_forces"):
        getattr(L, name).argtypes = [vp, vp]
    L.bh_get_bounds.argtypes = [vp, dp]
    L.bh_get_body_keys.argtypes = [vp, u32p]
    L.bh_get_sorted_order.argtypes = [vp, u32p]
    L.bh_get_tree_size.argtypes = [vp, i64p]
    L.bh_get_tree.argtypes = [vp, dp, C.c_int64, i64p]
    L.bh_dump_quadtree.argtypes = [vp, C.c_char_p]
    L.bh_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.bh_set_profiling.argtypes = [vp, C.c_int32]
    L.bh_get_timers.argtypes = [vp, C.POINTER(Timers)]
    L.bh_last_step_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.bh_direct_forces.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.bh_load_text.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64, dp, dp, dp]
    L.bh_append_positions_txt.argtypes = [C.c_char_p, dp, C.c_int64, C.c_double, C.c_int]
    L.bh_measure_fp32_peak.argtypes = [C.c_int32, dp, dp]
    L.bh_trajectory_begin.argtypes = [vp, C.c_char_p, C.c_int32]
    L.bh_trajectory_record.argtypes = [vp, C.c_double]
    L.bh_trajectory_end.argtypes = [vp]
    L.bh_generate.argtypes = [vp, C.c_int32, C.c_uint64]
    L.bh_generate_host.argtypes = [C.c_int32, C.c_uint64, C.c_int64, C.c_int64, C.c_int32, dp, dp, dp]
    L.bh_philox4x32_10.argtypes = [u32p, u32p, u32p]
    L.bh_write_init_files.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64, dp, dp, dp]
    for name in ABI_SYMBOLS:
        fn = getattr(L, name)
        if name in ("bh_default_params", "bh_philox4x32_10"):
            fn.restype = None
        elif name != "bh_last_error":
            fn.restype = C.c_int
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise BhError(rc, lib().bh_last_error().decode(errors="replace"))


def default_params(**over) -> Params:
    p = Params()
    lib().bh_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def shard_range(n_bodies: int, n_ranks: int, rank: int):
    lo, hi = C.c_int64(), C.c_int64()
    _check(lib().bh_shard_range(n_bodies, n_ranks, rank, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def _nccl_first_use():
    """libbh.so dlopens libnccl on first use and prefers a copy the process has already loaded.  A python host that
    also uses torch must have torch's bundled libnccl loaded FIRST: two libraries with the SONAME libnccl.so.2
    cannot coexist, and the system copy loaded first would break a later ``import torch``."""
    import importlib.util
    import sys
    if "torch" not in sys.modules and importlib.util.find_spec("torch") is not None:
        import torch  # noqa: F401


def nccl_unique_id() -> bytes:
    _nccl_first_use()
    buf = C.create_string_buffer(128)
    _check(lib().bh_nccl_unique_id(buf))
    return buf.raw


def measure_fp32_peak(device: int = -1):
    tf, mhz = C.c_double(), C.c_double()
    _check(lib().bh_measure_fp32_peak(device, C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


def generate_host(kind: str, n_bodies: int, seed: int = 12345, first: int = 0, round6: bool = False):
    """Seeded bodies [first, first + n_bodies) of the counter-based generators (csrc/generate.cu) on the host —
    the same values ``Simulation.generate`` writes on the device (up to the last bits of sqrt / sin / cos / pow).
    Returns (pos, vel, mass)."""
    pos = np.empty((n_bodies, 2)); vel = np.empty((n_bodies, 2)); mass = np.empty(n_bodies)
    dp = C.POINTER(C.c_double)
    _check(lib().bh_generate_host(GENERATOR_KINDS[kind], seed, first, n_bodies, int(round6), pos.ctypes.data_as(dp),
                                  vel.ctypes.data_as(dp), mass.ctypes.data_as(dp)))
    return pos, vel, mass


def append_positions_txt(path: str, pos, time: float, truncate: bool = False):
    """savePositions (project.cu:855-863): append one frame "time i x y \\n" per body to `path` (synchronous)."""
    pos = _f64(pos, (-1, 2))
    _check(lib().bh_append_positions_txt(path.encode(), pos.ctypes.data_as(C.POINTER(C.c_double)), pos.shape[0],
                                         float(time), 1 if truncate else 0))


def philox4x32_10(counter, key):
    c = (C.c_uint32 * 4)(*counter); k = (C.c_uint32 * 2)(*key); out = (C.c_uint32 * 4)()
    lib().bh_philox4x32_10(c, k, out)
    return list(out)


def write_init_files(directory: str, pos, vel, mass):
    """The reference's three initial-condition files (writers of project.cu:236-246, :268-281) through the C-ABI."""
    pos, vel, mass = _f64(pos, (-1, 2)), _f64(vel, (-1, 2)), _f64(mass, (-1,))
    if not (pos.shape[0] == vel.shape[0] == mass.shape[0]):
        raise ValueError("pos, vel and mass must describe the same number of bodies")
    os.makedirs(directory, exist_ok=True)
    dp = C.POINTER(C.c_double)
    _check(lib().bh_write_init_files(os.path.join(directory, "masses_init.txt").encode(),
                                     os.path.join(directory, "positions_init.txt").encode(),
                                     os.path.join(directory, "velocities_init.txt").encode(), mass.shape[0],
                                     mass.ctypes.data_as(dp), pos.ctypes.data_as(dp), vel.ctypes.data_as(dp)))


def load_text(directory: str, n_bodies: int):
    """loadSimulationDataFromText (project.cu:103-161) through the C-ABI."""
    mass = np.empty(n_bodies)
    pos = np.empty((n_bodies, 2))
    vel = np.empty((n_bodies, 2))
    dp = C.POINTER(C.c_double)
    _check(lib().bh_load_text(os.path.join(directory, "masses_init.txt").encode(),
                              os.path.join(directory, "positions_init.txt").encode(),
                              os.path.join(directory, "velocities_init.txt").encode(), n_bodies,
                              mass.ctypes.data_as(dp), pos.ctypes.data_as(dp), vel.ctypes.data_as(dp)))
    return pos, vel, mass


def _ptr(a):
    """Raw host pointer of a numpy array or a (pinned) torch CPU tensor."""
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def _f64(a, shape):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape)


class Simulation:
    """Host-side mirror of ``runSimulationGpu`` (project.cu:918-1024) over the C-ABI."""

    def __init__(self, n_bodies: int, fp64: bool = False, counters: bool = False, graph: bool = True,
                 exact_eps: bool = False, exact_leaves: bool = False, **over):
        flags = int(over.pop("flags", 0))
        if fp64:
            flags |= BH_FLAG_FP64_TRAVERSAL
        if counters:
            flags |= BH_FLAG_COUNTERS
        if not graph:
            flags |= BH_FLAG_NO_GRAPH
        if exact_eps:
            flags |= BH_FLAG_EXACT_EPS
        if exact_leaves:
            flags |= BH_FLAG_EXACT_LEAVES
        bpl = int(over.pop("bodies_per_lane", 0))
        self.params = default_params(n_bodies=n_bodies, flags=flags, **over)
        self.params.reserved[0] = bpl      # traversal tuning knob (include/bh.h)
        self.n = int(n_bodies)
        self._h = C.c_void_p()
        _check(lib().bh_create(C.byref(self.params), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().bh_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- state ----
    def attach_nccl(self, unique_id: bytes):
        _nccl_first_use()
        buf = C.create_string_buffer(unique_id, 128)
        _check(lib().bh_attach_nccl(self._h, buf))

    def comm_handle(self) -> bytes:
        """64-byte cudaIpc handle of this rank's peer-exchange buffer (all-gather, then attach_peers)."""
        buf = C.create_string_buffer(64)
        _check(lib().bh_comm_handle(self._h, buf))
        return buf.raw

    def attach_peers(self, handles):
        blob = b"".join(handles)
        buf = C.create_string_buffer(blob, len(blob))
        _check(lib().bh_attach_peers(self._h, buf, len(handles)))

    def set_bodies(self, pos, vel, mass):
        if not hasattr(pos, "data_ptr"):
            pos, vel, mass = _f64(pos, (self.n, 2)), _f64(vel, (self.n, 2)), _f64(mass, (self.n,))
        _check(lib().bh_set_bodies(self._h, _ptr(pos), _ptr(vel), _ptr(mass)))

    def set_positions(self, pos):
        if not hasattr(pos, "data_ptr"):
            pos = _f64(pos, (self.n, 2))
        _check(lib().bh_set_positions(self._h, _ptr(pos)))

    def set_velocities(self, vel):
        if not hasattr(vel, "data_ptr"):
            vel = _f64(vel, (self.n, 2))
        _check(lib().bh_set_velocities(self._h, _ptr(vel)))

    def generate(self, kind: str, seed: int = 12345):
        """Fill the context's bodies on the device with a seeded distribution (a rank of a multi-rank context
        generates only its slice); see generate_host for the host twin."""
        _check(lib().bh_generate(self._h, GENERATOR_KINDS[kind], seed))

    def snapshot(self):
        _check(lib().bh_snapshot(self._h))

    def restore(self):
        _check(lib().bh_restore(self._h))

    # ---- hot path ----
    def step(self, nsteps: int = 1):
        _check(lib().bh_step(self._h, nsteps))

    def step_host(self, pos, vel, mass, out_pos=None):
        """One step with host buffers (upload, step, download of the new positions), pipelined."""
        if not hasattr(pos, "data_ptr"):
            pos, vel, mass = _f64(pos, (self.n, 2)), _f64(vel, (self.n, 2)), _f64(mass, (self.n,))
        if out_pos is None:
            out_pos = np.empty((self.n, 2), dtype=np.float64)
        _check(lib().bh_step_host(self._h, _ptr(pos), _ptr(vel), _ptr(mass), _ptr(out_pos)))
        return out_pos

    def step_from_snapshot(self, nsteps: int = 1):
        _check(lib().bh_step_from_snapshot(self._h, nsteps))

    def build_tree(self):
        _check(lib().bh_build_tree(self._h))

    def compute_forces(self):
        _check(lib().bh_compute_forces(self._h))

    def integrate(self):
        _check(lib().bh_integrate(self._h))

    def synchronize(self):
        _check(lib().bh_synchronize(self._h))

    # ---- results ----
    def _get2(self, fn, out=None):
        if out is None:
            out = np.empty((self.n, 2), dtype=np.float64)
        _check(fn(self._h, _ptr(out)))
        return out

    def positions(self, out=None):
        return self._get2(lib().bh_get_positions, out)

    def velocities(self, out=None):
        return self._get2(lib().bh_get_velocities, out)

    def accelerations(self, out=None):
        return self._get2(lib().bh_get_accelerations, out)

    def forces(self, out=None):
        return self._get2(lib().bh_get_forces, out)

    def bounds(self):
        out = np.empty(4)
        _check(lib().bh_get_bounds(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def body_keys(self):
        out = np.empty(self.n, dtype=np.uint32)
        _check(lib().bh_get_body_keys(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def sorted_order(self):
        out = np.empty(self.n, dtype=np.uint32)
        _check(lib().bh_get_sorted_order(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def tree_size(self) -> int:
        v = C.c_int64()
        _check(lib().bh_get_tree_size(self._h, C.byref(v)))
        return v.value

    def tree(self) -> np.ndarray:
        """Canonical node table (rows of 10 doubles), see bh_get_tree in include/bh.h."""
        n = self.tree_size()
        out = np.empty((n, 10), dtype=np.float64)
        got = C.c_int64()
        _check(lib().bh_get_tree(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), n, C.byref(got)))
        if got.value != n:
            raise BhError(-1, f"tree walk produced {got.value} rows, device counted {n}")
        return out

    def dump_quadtree(self, path: str):
        _check(lib().bh_dump_quadtree(self._h, path.encode()))

    def counters(self) -> dict:
        c = Counters()
        _check(lib().bh_get_counters(self._h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_ if k != "reserved"}

    def set_profiling(self, on: bool):
        _check(lib().bh_set_profiling(self._h, 1 if on else 0))

    def timers(self) -> dict:
        t = Timers()
        _check(lib().bh_get_timers(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timers._fields_}

    def reset_timers(self):
        _check(lib().bh_reset_timers(self._h))

    def last_step_ms(self) -> float:
        ms = C.c_float()
        _check(lib().bh_last_step_ms(self._h, C.byref(ms)))
        return ms.value

    # ---- trajectory output (positions.txt of plot_2d.py), asynchronous ----
    def trajectory_begin(self, path: str, stride: int = 1):
        _check(lib().bh_trajectory_begin(self._h, path.encode(), stride))

    def trajectory_record(self, time: float):
        _check(lib().bh_trajectory_record(self._h, C.c_double(time)))

    def trajectory_end(self):
        _check(lib().bh_trajectory_end(self._h))

    def direct_forces(self, want_output: bool = True):
        ms = C.c_float()
        out = np.empty((self.n, 2)) if want_output else None
        _check(lib().bh_direct_forces(self._h, _ptr(out) if want_output else None, C.byref(ms)))
        return out, ms.value


def run_simulation(pos, vel, mass, n_steps: int, workdir: str = ".", positions_txt: bool = False, **params):
    """``runSimulationGpu`` as the reference's main drives it (project.cu:918-1024): writes
    quadtree_init_gpu.txt at step 0 and quadtree_final_gpu.txt at the last step (only when
    n_steps >= 2: project.cu:962-965), optionally the trajectory file of the CPU programs."""
    n = np.asarray(mass).shape[0]
    with Simulation(n, **params) as sim:
        sim.set_bodies(pos, vel, mass)
        traj = os.path.join(workdir, "positions.txt")
        dp = C.POINTER(C.c_double)
        if positions_txt:
            p0 = _f64(pos, (n, 2))
            _check(lib().bh_append_positions_txt(traj.encode(), p0.ctypes.data_as(dp), n, 0.0, 1))
        open(os.path.join(workdir, "quadtree_final_gpu.txt"), "w").close()   # reference opens both files
        t = 0.0
        for s in range(n_steps):
            t += sim.params.dt
            first, last = s == 0, (s == n_steps - 1 and s != 0)
            if first or last or positions_txt:
                sim.build_tree()
                if first:
                    sim.dump_quadtree(os.path.join(workdir, "quadtree_init_gpu.txt"))
                elif last:
                    sim.dump_quadtree(os.path.join(workdir, "quadtree_final_gpu.txt"))
            sim.step(1)
            if positions_txt:
                p = sim.positions()
                _check(lib().bh_append_positions_txt(traj.encode(), p.ctypes.data_as(dp), n, t, 0))
        return sim.positions(), sim.velocities()
