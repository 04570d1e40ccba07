"""Per-step and per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file x.csv`):
    python tools/launch_list_summary.py gpurun_out/r2p_launches_bench.csv ["title line"]
One timed step = the kernels from one bounds_kernel to the next; the LAST complete step of the command is printed
(the steady state: earlier ones include first-touch effects), then every kernel of the command with count and total."""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
launches = []
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]]
    name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("bh::<unnamed>::", "").strip()
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = val / 1000.0 if unit in ("ns", "nsecond") else val * 1000.0 if unit in ("ms", "msecond") else val
    launches.append((name, us))
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print(f"# {len(launches)} launches in the whole command; gpu__time_duration.sum per launch, --clock-control none: cold-cache, serialised")
starts = [i for i, (n, _) in enumerate(launches) if n == "bounds_kernel"]
steps = [launches[a:b] for a, b in zip(starts, starts[1:]) if any("traverse" in n for n, _ in launches[a:b])]
# the timed step of bench.py is the one whose traversal kernel carries the fused integrator (no separate integrate_kernel)
full = [s for s in steps if any(n.startswith("traverse_f32_list_kernel<1") for n, _ in s)] or steps
full = [s for s in full if len(s) == min(len(x) for x in full)]          # (other legs append getters / integrators)
if full:
    s = sorted(full, key=lambda st: sum(u for _, u in st))[len(full) // 2]   # the median step
    tot = sum(u for _, u in s)
    print(f"# one timed step = the {len(s)} kernels from one bounds_kernel to the next (the median such step of the command):")
    for n, u in s:
        print(f"  {n:<40s} {u:8.2f} us  {100 * u / tot:5.1f} %")
    print(f"  {'step total (sum of kernels)':<40s} {tot:8.2f} us")
agg = OrderedDict()
for n, u in launches:
    c, t = agg.get(n, (0, 0.0)); agg[n] = (c + 1, t + u)
print("# all launches of the command, by kernel:")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:<60s} x{c:4d} {t:11.1f} us")
