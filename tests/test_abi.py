"""CPU suite: the C-ABI library loads, exports every symbol include/bh.h declares, its host-only
entry points behave like the reference's, and it fails loudly without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gpu_nbody_simulation_b200 as bh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bh.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = bh.lib()
    names = header_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(L, name), f"libbh.so lacks {name}"
    assert sorted(bh.ABI_SYMBOLS) == names, "python binding and header disagree"
    assert L.bh_abi_version() == 1


def test_params_struct_matches_header_defaults():
    p = bh.default_params()
    assert (p.n_bodies, p.G, p.dt, p.theta) == (40000, 6.67e-11, 1.0, 0.5)      # project.cu:1-3, :27, :29, :60
    assert (p.dist_eps, p.mass_eps, p.pad_frac, p.pad_fallback) == (1e-15, 1e-15, 0.1, 1e-6)
    assert (p.max_depth, p.n_ranks, p.rank) == (10, 1, 0)                        # project.cu:61
    assert C.sizeof(bh.Params) == 8 * 8 + 4 * 6 + 16


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bh.BhError) as e:
        bh.Simulation(1000)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_shard_range_partitions():
    for n in (0, 1, 7, 40000, 1_000_003, 16_777_216):
        for r in (1, 2, 3, 8):
            prev = 0
            for k in range(r):
                lo, hi = bh.shard_range(n, r, k)
                assert lo == prev and hi >= lo
                assert abs((hi - lo) - n / r) < 1.0
                prev = hi
            assert prev == n
    with pytest.raises(bh.BhError):
        bh.shard_range(10, 2, 2)


def test_load_text_reads_reference_format(tmp_path):
    pos, vel, mass = bh.initial_conditions.uniform_square(300, seed=3)
    bh.initial_conditions.write_init_files(str(tmp_path), pos, vel, mass)
    p, v, m = bh.load_text(str(tmp_path), 257)           # first N lines only (project.cu:121, :137)
    assert np.array_equal(p, pos[:257]) and np.array_equal(v, vel[:257]) and np.array_equal(m, mass[:257])
    with pytest.raises(bh.BhError) as e:                  # too few lines -> error (project.cu:122-124)
        bh.load_text(str(tmp_path), 301)
    assert "Not enough" in str(e.value)
    with pytest.raises(bh.BhError) as e:
        bh.load_text(str(tmp_path / "missing"), 1)
    assert "Failed to open file" in str(e.value)


def test_positions_txt_format(tmp_path):
    path = str(tmp_path / "positions.txt")
    pos = np.array([[0.0558754321, -0.0739181], [1.5, 2.0]])
    dp = C.POINTER(C.c_double)
    assert bh.lib().bh_append_positions_txt(path.encode(), pos.ctypes.data_as(dp), 2, 0.0, 1) == 0
    assert bh.lib().bh_append_positions_txt(path.encode(), pos.ctypes.data_as(dp), 2, 1.0, 0) == 0
    lines = open(path).read().split("\n")
    # std::to_string formatting, trailing space (project.cu:857-861); plot_2d.py reads 4 floats per line
    assert lines[0] == "0.000000 0 0.055875 -0.073918 "
    assert lines[3] == "1.000000 1 1.500000 2.000000 "
    data = np.loadtxt(path)
    assert data.shape == (4, 4)


def test_initial_condition_generators_are_seeded():
    a = bh.initial_conditions.uniform_disk(5000, seed=12345)
    b = bh.initial_conditions.uniform_disk(5000, seed=12345)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    pos, vel, mass = a
    assert np.all(np.hypot(pos[:, 0], pos[:, 1]) <= 0.1 + 1e-6)
    assert mass.min() >= 0.1 - 1e-6 and mass.max() <= 0.5 + 1e-6 and np.abs(vel).max() <= 1e-4 + 1e-9
    p2, _, _ = bh.initial_conditions.plummer_2d(5000, seed=1)
    assert np.all(np.hypot(p2[:, 0], p2[:, 1]) <= 0.1 + 1e-6)
    # %.6g round trip: values survive the reference's text writers unchanged
    assert np.array_equal(np.char.mod("%.6g", pos.ravel()).astype(float), pos.ravel())


def test_header_is_plain_c_and_struct_layouts_match_the_python_mirror(tmp_path):
    """include/bh.h must compile as C (the boundary is a C-ABI, no C++ / torch types) and the ctypes mirrors
    must have the C compiler's sizes and field offsets."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    fields = {"bh_params": [f[0] for f in bh.Params._fields_], "bh_counters": [f[0] for f in bh.Counters._fields_],
              "bh_timers": [f[0] for f in bh.Timers._fields_]}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "bh.h"', 'int main(void) {']
    for st, fl in fields.items():
        lines.append(f'  printf("{st} %zu\\n", sizeof({st}));')
        for f in fl:
            lines.append(f'  printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    lines += ['  printf("flags %u %u %u %u %u\\n", BH_FLAG_FP64_TRAVERSAL, BH_FLAG_COUNTERS, BH_FLAG_NO_GRAPH, BH_FLAG_EXACT_EPS, BH_FLAG_EXACT_LEAVES);',
              '  printf("gen %d %d %d\\n", BH_GEN_UNIFORM_SQUARE, BH_GEN_UNIFORM_DISK, BH_GEN_PLUMMER_2D);',
              '  return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = dict(l.split(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for st, cls in (("bh_params", bh.Params), ("bh_counters", bh.Counters), ("bh_timers", bh.Timers)):
        assert int(out[st]) == C.sizeof(cls), st
        for f in fields[st]:
            assert int(out[f"{st}.{f}"]) == getattr(cls, f).offset, f"{st}.{f}"
    assert out["flags"].split() == [str(v) for v in (bh.BH_FLAG_FP64_TRAVERSAL, bh.BH_FLAG_COUNTERS, bh.BH_FLAG_NO_GRAPH,
                                                      bh.BH_FLAG_EXACT_EPS, bh.BH_FLAG_EXACT_LEAVES)]
    assert out["gen"].split() == [str(bh.GENERATOR_KINDS[k]) for k in ("uniform_square", "uniform_disk", "plummer_2d")]


def test_integration_stub_compiles_and_links(tmp_path):
    """The reference-side binding shown in INTEGRATION.md (the replacement body of runSimulationGpu) must stay
    valid: compile it with the reference's own typedefs (project.cu:38-43) and link it against libbh.so."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = re.search(r"```cpp\n(.*?)```", md, flags=re.S).group(1)
    assert "runSimulationGpu" in stub and "bh_step" in stub
    src = tmp_path / "stub.cpp"
    src.write_text("#include <array>\n#include <cstdlib>\n#include <iostream>\n#define N_BODIES 1000\n#define N_SIMULATIONS 3\n"
                   "using Vector = std::array<double, 2>;\nusing Positions = std::array<Vector, N_BODIES>;\n"
                   "using Velocities = Positions;\nusing Masses = std::array<double, N_BODIES>;\n"
                   "long long gpu_parallel_duration = 0;\n" + stub +
                   "\nint main() { static Masses m; static Positions p; static Velocities v; if (m[0] > 1) runSimulationGpu(m, p, v); return 0; }\n")
    pkg = os.path.join(ROOT, "gpu_nbody_simulation_b200")
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-L", pkg, "-lbh",
                           "-Wl,--unresolved-symbols=ignore-in-shared-libs", "-o", str(tmp_path / "stub")])


def test_positions_txt_equals_the_live_references_savePositions(tmp_path):
    """bh_append_positions_txt (host code of libbh.so, no GPU needed) against the reference's own savePositions
    (project.cu:855-863) run live through oracle/_ref: the trajectory file plot_2d.py reads, byte for byte, for the
    positions of a 2-step reference run (t = 0, 1, 2)."""
    import oracle
    n = 2048
    if not oracle.ref_available(n):
        pytest.skip(f"oracle/_ref/ref_harness_N{n} not built here")
    rng = np.random.default_rng(12)
    pos = rng.uniform(-0.1, 0.1, size=(n, 2))
    vel = rng.uniform(-1e-4, 1e-4, size=(n, 2))
    mass = np.power(10.0, rng.uniform(-1.0, np.log10(0.5), size=n))
    ref_file = str(tmp_path / "positions_ref.txt")
    recs, _ = oracle.run_ref(pos, vel, mass, steps=2, positions_txt=ref_file)
    mine = str(tmp_path / "positions.txt").encode()
    dp = C.POINTER(C.c_double)
    L = bh.lib()
    assert L.bh_append_positions_txt(mine, np.ascontiguousarray(pos).ctypes.data_as(dp), n, 0.0, 1) == 0
    for s in range(2):
        p = np.ascontiguousarray(recs[("pos", s)])
        assert L.bh_append_positions_txt(mine, p.ctypes.data_as(dp), n, float(s + 1), 0) == 0
    assert open(mine.decode(), "rb").read() == open(ref_file, "rb").read()


def test_load_text_equals_the_live_references_loader(tmp_path):
    """bh_load_text against the reference's own loadSimulationDataFromText (project.cu:103-161) run live: same doubles
    from the same three files (written by bh_write_init_files), extra lines ignored, and the same refusal when a
    file is too short."""
    import oracle
    n = 2048
    if not oracle.ref_available(n):
        pytest.skip(f"oracle/_ref/ref_harness_N{n} not built here")
    pos, vel, mass = bh.generate_host("plummer_2d", n + 7, seed=21)            # 7 lines more than N: first N are used
    d = str(tmp_path / "ic")
    bh.write_init_files(d, pos, vel, mass)
    p_ref, v_ref, m_ref = oracle.run_ref_loader(d, n)
    p, v, m = bh.load_text(d, n)
    assert np.array_equal(p, p_ref) and np.array_equal(v, v_ref) and np.array_equal(m, m_ref)
    with pytest.raises(ValueError):
        bh.write_init_files(str(tmp_path / "bad"), pos[:n - 1], vel, mass)
    short = str(tmp_path / "short")
    bh.write_init_files(short, pos, vel, mass)
    lines = open(os.path.join(short, "positions_init.txt")).read().splitlines()
    open(os.path.join(short, "positions_init.txt"), "w").write("\n".join(lines[:n - 1]) + "\n")   # one line short
    with pytest.raises(RuntimeError) as ref_err:
        oracle.run_ref_loader(short, n)
    with pytest.raises(bh.BhError) as my_err:
        bh.load_text(short, n)
    assert "Not enough" in str(ref_err.value) and "Not enough" in str(my_err.value)
