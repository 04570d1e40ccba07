"""CPU tests of bench.py's host-side helpers (no GPU): clock-sample parsing, NUMA pinning fall-back,
the reference arm's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class _FakeProc:
    def terminate(self):
        pass


class _FakeThread:
    def join(self, timeout=None):
        pass


def test_clock_sampler_parses_only_samples_after_the_mark():
    s = bench.ClockSampler("0,1")
    s.proc, s.t = _FakeProc(), _FakeThread()
    s.lines = ["0, 1200, 1965, 300.1, 0x0, Not Active, Not Active, Not Active, Not Active\n"]   # before the mark
    s.mark()
    s.lines += ["0, 1965, 1965, 900.5, 0x4, Not Active, Not Active, Not Active, Active\n",
                "1, 1950, 1965, 880.0, 0x0, Not Active, Not Active, Not Active, Not Active\n",
                "garbage line\n",
                "1, [N/A], 1965, 880.0, 0x0, Not Active, Not Active, Not Active, Not Active\n"]
    out = s.stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]


def test_clock_sampler_without_nvidia_smi():
    s = bench.ClockSampler("0")
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]


def test_numa_pinning_is_best_effort():
    before = os.sched_getaffinity(0)
    assert bench.pin_to_gpu_numa_node(0) is None or isinstance(bench.pin_to_gpu_numa_node(0), int)
    os.sched_setaffinity(0, before)


def test_reference_arm_prints_the_contract_line():
    # 65 536-body sample keeps this to a few seconds; under torchrun only rank 0 prints
    env = dict(os.environ, RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
    line = bench.cpu_reference_run(65_536, 1, 0, budget_s=5.0)[1]
    assert line["kind"] in ("reference", "port") and line["cores"] == 1 and line["value"] > 0
    assert line["n_bodies"] == 65_536 and line["steps_executed"] >= 1     # the body count timed is a structured field
    json.dumps(line)


def test_both_arms_name_the_same_workload():
    """The driver compares the two arms' config: the reference arm must run (and say it runs) the N = 1M workload."""
    assert bench.workload_string(1_000_000, 1) == bench.workload_string(bench.BODIES_PER_GPU, 1, "disk", 10)
    assert "N=1000000" in bench.workload_string(1_000_000, 1)
    import inspect
    src = inspect.getsource(bench.run_reference)
    assert "workload_string(n, 1)" in src and "cpu_reference_run(n," in src and "n = BODIES_PER_GPU" in src


def test_bracket_plan_and_workload_generator():
    assert bench.bracket_plan(20) == 50 and bench.bracket_plan(200) == 5 and bench.bracket_plan(1000) == 5
    import numpy as np
    pos, vel, mass = bench.make_workload(1000)
    pos2, _, _ = bench.make_workload(2000)
    assert np.array_equal(pos, pos2[:1000])                      # body i is a pure function of (seed, i)
    assert (np.hypot(pos[:, 0], pos[:, 1]) <= 0.1 + 1e-12).all() and mass.min() >= 0.1 and mass.max() <= 0.5


def test_shell_scripts_parse():
    import glob
    scripts = glob.glob(os.path.join(ROOT, "scripts", "*.sh")) + glob.glob(os.path.join(ROOT, "tools", "*.sh")) + \
        [os.path.join(ROOT, "oracle", "build_ref.sh")]
    assert len(scripts) >= 6
    for s in scripts:
        assert subprocess.run(["bash", "-n", s]).returncode == 0, s


def _bench_line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(next(ln for ln in f if ln.startswith("{")))


def test_committed_bench_lines_obey_the_contract():
    """The bench lines committed under profiles/ (printed by bench.py on B200s) carry every key of the measurement
    contract and their numbers are consistent with each other — a guard against the docs and the evidence drifting
    apart, and against a bench.py edit that drops a key."""
    d = _bench_line("r02_bench_default.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "accuracy"):
        assert k in d, k
    assert d["metric"] == bench.METRIC and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    n = d["config"]["n_bodies"]
    assert n == bench.BODIES_PER_GPU and "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - n / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]               # body-steps/s = N / step time
    r = d["roofline"]
    assert r["bound"] == "fp32" and abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-9
    assert abs(r["achieved"] - r["interactions_per_step"] * r["flop_per_interaction"] / (r["kernel_us"] * 1e-6) / 1e12) <= 1e-6 * r["achieved"]
    assert r["kernel_us"] * 1e-3 <= d["ms_per_step"] * 1.1
    # (the line was printed with the ncu capture of the previous build of the same kernel: 88.5 MB; the final one: 87.8 MB)
    assert 0.95 < r["traffic"] / bench.TRAVERSE_DRAM_BYTES_NCU < 1.05 and r["traffic"] >= 72 * n
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 40 * n and e["d2h_bytes_per_step"] == 16 * n and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] == 1 and 0 < c["value"] < 1e-3 * d["value"]
    assert d["accuracy"]["force_rel_rms_vs_reference_tree"] <= d["accuracy"]["bar"] == 1e-5
    assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for name, gpus in (("r02_bench_g2.json", 2), ("r02_bench_g8.json", 8)):
        m = _bench_line(name)
        assert m["n_gpus"] == gpus and m["scaling"] == "weak" and m["config"]["n_bodies"] == gpus * bench.BODIES_PER_GPU
        assert m["accuracy"]["force_rel_rms_vs_reference_tree"] <= 1e-5 and m["roofline"]["frac"] > 0
        s = m["strong"]
        assert s and s["speedup_vs_1gpu"] > 1 and s["accuracy"]["force_rel_rms_vs_reference_tree"] <= 1e-5
