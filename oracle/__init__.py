"""TEST INFRASTRUCTURE — ctypes front end of the CPU oracle (oracle/bh_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package, and only as the checker / the timed CPU baseline.  The product package
(gpu_nbody_simulation_b200) never imports it.

The C file restates ``implementation/project.cu`` of the reference (file:line citations are in
the C source); ``ref_harness`` below drives the reference's OWN compiled CPU functions when
``oracle/_ref/ref_harness_N<N>`` has been built (oracle/build_ref.sh).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libbh_oracle.so")
_lib = None


class Params(C.Structure):
    """Runtime copy of the reference's source-level constants (project.cu:27-35, :60-61)."""

    _fields_ = [
        ("G", C.c_double), ("dt", C.c_double), ("theta", C.c_double), ("dist_eps", C.c_double),
        ("mass_eps", C.c_double), ("pad_frac", C.c_double), ("pad_fallback", C.c_double),
        ("max_depth", C.c_int32), ("_pad", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/bh_oracle.c (gcc) if needed; returns the .so path."""
    src = os.path.join(_HERE, "bh_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "_build/libbh_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        i64p = C.POINTER(C.c_int64)
        pp = C.POINTER(Params)
        L.bho_default_params.argtypes = [pp]
        L.bho_root_bounds.argtypes = [dp, C.c_int64, pp, dp]
        L.bho_build_tree.argtypes = [dp, dp, C.c_int64, pp]
        L.bho_build_tree.restype = C.c_void_p
        L.bho_tree_size.argtypes = [C.c_void_p]
        L.bho_tree_size.restype = C.c_int64
        L.bho_tree_nodes.argtypes = [C.c_void_p]
        L.bho_tree_nodes.restype = dp
        L.bho_tree_free.argtypes = [C.c_void_p]
        L.bho_compute_forces.argtypes = [C.c_void_p, dp, dp, C.c_int64, pp, C.c_int64, C.c_int64,
                                         C.c_int64, C.c_int, dp, i64p]
        L.bho_compute_forces_exact_leaves.argtypes = L.bho_compute_forces.argtypes
        L.bho_update.argtypes = [dp, dp, dp, dp, dp, C.c_int64, C.c_double]
        L.bho_step.argtypes = [dp, dp, dp, dp, dp, C.c_int64, pp, C.c_int, i64p]
        L.bho_step.restype = C.c_int64
        L.bho_body_keys.argtypes = [dp, C.c_int64, dp, C.c_int, C.POINTER(C.c_uint32)]
        L.bho_canonical_table.argtypes = [C.c_void_p, dp, C.c_int64]
        L.bho_canonical_table.restype = C.c_int64
        L.bho_dump_quadtree.argtypes = [C.c_void_p, dp, C.c_char_p]
        L.bho_dump_quadtree.restype = C.c_int
        L.bho_direct_forces.argtypes = [dp, dp, C.c_int64, C.c_double, C.c_int64, C.c_int64, C.c_int, dp]
        L.bho_max_threads.restype = C.c_int
        _lib = L
    return _lib


def default_params(**over) -> Params:
    p = Params()
    lib().bho_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


COUNTER_NAMES = ("visits", "interactions", "opens", "zero_skips", "self_skips", "max_stack")


class Tree:
    """The reference's node array: rows of 12 doubles (project.cu:46-58)."""

    def __init__(self, pos, mass, params: Params | None = None):
        self.params = params or default_params()
        self.pos = _f64(pos, (-1, 2))
        self.mass = _f64(mass, (-1,))
        self.n = self.mass.shape[0]
        self._h = lib().bho_build_tree(_dp(self.pos), _dp(self.mass), self.n, C.byref(self.params))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bho_tree_free(self._h)
            self._h = None

    @property
    def size(self) -> int:
        return lib().bho_tree_size(self._h)

    def nodes(self) -> np.ndarray:
        n = self.size
        buf = lib().bho_tree_nodes(self._h)
        return np.ctypeslib.as_array(buf, shape=(n, 12)).copy()

    def canonical(self) -> np.ndarray:
        """DFS pre-order rows [depth, xmin, xmax, ymin, ymax, mass, comx, comy, occupant, internal]."""
        out = np.empty((self.size, 10), dtype=np.float64)
        got = lib().bho_canonical_table(self._h, _dp(out), out.shape[0])
        assert got == out.shape[0]
        return out

    def forces(self, i0=0, i1=None, stride=1, nthreads=1, pos=None, mass=None):
        pos = self.pos if pos is None else _f64(pos, (-1, 2))
        mass = self.mass if mass is None else _f64(mass, (-1,))
        i1 = self.n if i1 is None else i1
        f = np.zeros((self.n, 2), dtype=np.float64)
        cnt = (C.c_int64 * 6)()
        lib().bho_compute_forces(self._h, _dp(pos), _dp(mass), self.n, C.byref(self.params), i0, i1, stride,
                                 nthreads, _dp(f), cnt)
        return f, dict(zip(COUNTER_NAMES, list(cnt)))

    def forces_exact_leaves(self, i0=0, i1=None, stride=1, nthreads=1):
        """EXTENSION (not reference behaviour): multi-body cap-level leaves are applied
        as exact pair sums over their bodies, self excluded; everything else as forces()."""
        i1 = self.n if i1 is None else i1
        f = np.zeros((self.n, 2), dtype=np.float64)
        cnt = (C.c_int64 * 6)()
        lib().bho_compute_forces_exact_leaves(self._h, _dp(self.pos), _dp(self.mass), self.n, C.byref(self.params),
                                              i0, i1, stride, nthreads, _dp(f), cnt)
        return f, dict(zip(COUNTER_NAMES, list(cnt)))

    def dump(self, path: str):
        rc = lib().bho_dump_quadtree(self._h, _dp(self.pos), path.encode())
        if rc:
            raise OSError(path)


def root_bounds(pos, params: Params | None = None) -> np.ndarray:
    params = params or default_params()
    pos = _f64(pos, (-1, 2))
    out = np.empty(4)
    lib().bho_root_bounds(_dp(pos), pos.shape[0], C.byref(params), _dp(out))
    return out


def body_keys(pos, bounds, max_depth=10) -> np.ndarray:
    pos = _f64(pos, (-1, 2))
    b = _f64(bounds, (4,))
    keys = np.empty(pos.shape[0], dtype=np.uint32)
    lib().bho_body_keys(_dp(pos), pos.shape[0], _dp(b), max_depth, keys.ctypes.data_as(C.POINTER(C.c_uint32)))
    return keys


def step(pos, vel, mass, params: Params | None = None, nthreads=1):
    """One loop iteration of runSimulationCpu (project.cu:883-910), in place on copies.
    Returns dict(pos, vel, acc, forces, nodes, counters)."""
    params = params or default_params()
    pos = _f64(pos, (-1, 2)).copy()
    vel = _f64(vel, (-1, 2)).copy()
    mass = _f64(mass, (-1,))
    n = mass.shape[0]
    acc = np.zeros((n, 2))
    f = np.zeros((n, 2))
    cnt = (C.c_int64 * 6)()
    nodes = lib().bho_step(_dp(pos), _dp(vel), _dp(mass), _dp(acc), _dp(f), n, C.byref(params), nthreads, cnt)
    return dict(pos=pos, vel=vel, acc=acc, forces=f, nodes=int(nodes), counters=dict(zip(COUNTER_NAMES, list(cnt))))


def update(forces, mass, vel, pos, dt):
    forces = _f64(forces, (-1, 2))
    mass = _f64(mass, (-1,))
    vel = _f64(vel, (-1, 2)).copy()
    pos = _f64(pos, (-1, 2)).copy()
    acc = np.zeros_like(pos)
    lib().bho_update(_dp(forces), _dp(mass), _dp(acc), _dp(vel), _dp(pos), mass.shape[0], dt)
    return acc, vel, pos


def direct_forces(pos, mass, G=6.67e-11, i0=0, i1=None, nthreads=1):
    pos = _f64(pos, (-1, 2))
    mass = _f64(mass, (-1,))
    n = mass.shape[0]
    i1 = n if i1 is None else i1
    f = np.zeros((n, 2))
    lib().bho_direct_forces(_dp(pos), _dp(mass), n, G, i0, i1, nthreads, _dp(f))
    return f


def max_threads() -> int:
    return int(lib().bho_max_threads())


# ----------------------------------------------------------------------------------------------
# The reference's own compiled CPU path (oracle/_ref), when it has been built for this N.
# ----------------------------------------------------------------------------------------------
def ref_harness_path(n: int) -> str:
    return os.path.join(_HERE, "_ref", f"ref_harness_N{n}")


def ref_available(n: int) -> bool:
    return os.access(ref_harness_path(n), os.X_OK)


def write_bodies_bin(path, pos, vel, mass):
    pos = _f64(pos, (-1, 2)); vel = _f64(vel, (-1, 2)); mass = _f64(mass, (-1,))
    with open(path, "wb") as f:
        f.write(np.uint64(mass.shape[0]).tobytes())
        f.write(mass.tobytes()); f.write(pos.tobytes()); f.write(vel.tobytes())


def read_dump(path):
    """Parse ref_harness's record stream -> {(name, step): ndarray}."""
    out = {}
    with open(path, "rb") as f:
        while True:
            tag = f.read(16)
            if len(tag) < 16:
                break
            step, n = np.frombuffer(f.read(16), dtype=np.uint64)
            data = np.frombuffer(f.read(int(n) * 8), dtype=np.float64).copy()
            out[(tag.rstrip(b"\0").decode(), int(step))] = data
    return out


def run_ref(pos, vel, mass, steps=1, dump="tree,forces,state", dump_steps="all", quadtree_txt=None,
            reset_each_step=False, keep_dump=True, positions_txt=None):
    """Run the reference's own CPU functions (unmodified project.cu) on these bodies.
    Returns (records, timings) where timings is the list of per-step JSON dicts."""
    mass = _f64(mass, (-1,))
    n = mass.shape[0]
    exe = ref_harness_path(n)
    if not os.access(exe, os.X_OK):
        raise FileNotFoundError(f"{exe} not built (oracle/build_ref.sh {n})")
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "bodies.bin")
        outp = os.path.join(td, "dump.bin")
        write_bodies_bin(inp, pos, vel, mass)
        cmd = [exe, "--in", inp, "--steps", str(steps), "--dump", dump, "--dump-steps", dump_steps]
        if keep_dump:
            cmd += ["--out", outp]
        if quadtree_txt:
            cmd += ["--quadtree-txt", quadtree_txt]
        if reset_each_step:
            cmd += ["--reset-each-step"]
        if positions_txt:
            cmd += ["--positions-txt", positions_txt]
        res = subprocess.run(cmd, check=True, capture_output=True, text=True)
        timings = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
        recs = read_dump(outp) if keep_dump else {}
    return recs, timings


# ----------------------------------------------------------------------------------------------
# The reference's own GPU program path (runSimulationGpu, unmodified project.cu compiled for
# sm_100a by oracle/build_ref.sh gpu <N> <steps>) — the "reference project.cu on one B200" baseline.
# ----------------------------------------------------------------------------------------------
def ref_gpu_path(n: int, steps: int) -> str:
    return os.path.join(_HERE, "_ref", f"ref_gpu_N{n}_S{steps}")


def ref_gpu_available(n: int, steps: int) -> bool:
    return os.access(ref_gpu_path(n, steps), os.X_OK)


def run_ref_gpu(pos, vel, mass, steps=1, calls=2, want_positions=False, device=None, timeout=600):
    """Run the reference's runSimulationGpu (`steps` = its compile-time N_SIMULATIONS) `calls` times,
    every call from these initial bodies.  Returns (per-call JSON dicts, positions after the last
    call or None).  Call 0 includes the CUDA context creation: treat it as warm-up."""
    mass = _f64(mass, (-1,))
    n = mass.shape[0]
    exe = ref_gpu_path(n, steps)
    if not os.access(exe, os.X_OK):
        raise FileNotFoundError(f"{exe} not built (oracle/build_ref.sh gpu {n} {steps})")
    env = dict(os.environ)
    if device is not None:
        vis = [v for v in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        env["CUDA_VISIBLE_DEVICES"] = vis[device] if device < len(vis) else str(device)
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "bodies.bin")
        outp = os.path.join(td, "positions.bin")
        write_bodies_bin(inp, pos, vel, mass)
        cmd = [exe, "--in", inp, "--calls", str(calls)] + (["--out", outp] if want_positions else [])
        res = subprocess.run(cmd, cwd=td, env=env, capture_output=True, text=True, timeout=timeout)
        if res.returncode != 0:
            raise RuntimeError(f"{os.path.basename(exe)} exited {res.returncode}: {res.stderr.strip()[-300:]}")
        calls_out = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
        p = np.fromfile(outp, dtype=np.float64).reshape(n, 2) if want_positions else None
    return calls_out, p


def run_ref_loader(directory: str, n: int):
    """The reference's own loadSimulationDataFromText (project.cu:103-161) on DIR/masses_init.txt, positions_init.txt,
    velocities_init.txt (first n lines).  Returns (pos, vel, mass); raises RuntimeError with the reference's exception
    text when it throws."""
    exe = ref_harness_path(n)
    if not os.access(exe, os.X_OK):
        raise FileNotFoundError(f"{exe} not built (oracle/build_ref.sh {n})")
    with tempfile.TemporaryDirectory() as td:
        outp = os.path.join(td, "dump.bin")
        res = subprocess.run([exe, "--load-text", directory, "--steps", "0", "--out", outp, "--dump", ""],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(res.stderr.strip())
        recs = read_dump(outp)
    return recs[("loaded_pos", 0)].reshape(n, 2), recs[("loaded_vel", 0)].reshape(n, 2), recs[("loaded_mass", 0)]
