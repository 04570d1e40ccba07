"""Execution-count regions of a kernel's SASS (consecutive instructions with the same execution count) from an .ncu-rep:
    python tools/sass_regions.py rep.ncu-rep [min_share]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]; ix = {n: i for i, n in enumerate(h)}
body = [r for r in rows[hdr + 1:] if len(r) == len(h)]
groups = []
for r in body:
    c = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    if groups and groups[-1][0] == c:
        groups[-1][1] += 1; groups[-1][2] += s
    else:
        groups.append([c, 1, s, r[ix["Source"]].strip()])
tot = sum(g[0] * g[1] for g in groups); tots = sum(g[2] for g in groups)
print(f"total warp instructions {tot}, samples {tots}")
for g in groups:
    if g[0] * g[1] > min_share * tot:
        print(f"exec {g[0]:9d} x {g[1]:3d} instr = {100 * g[0] * g[1] / tot:5.1f}% inst, {100 * g[2] / tots:5.1f}% samples; first: {g[3][:70]}")
