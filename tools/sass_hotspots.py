"""Per-opcode and per-region instruction / stall summary of one kernel from an .ncu-rep source page.
    python tools/sass_hotspots.py rep.ncu-rep [--list]   (needs `ncu`; reads the report on the CPU box)"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
ix = {n: i for i, n in enumerate(h)}
body = [r for r in rows[hdr + 1:] if len(r) == len(h)]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
tot_samp = sum(int(r[ix["# Samples"]]) for r in body)
print(f"kernel: {rows[0][1][:100]}\ninstructions executed {tot_inst}, stall samples {tot_samp}, SASS lines {len(body)}")
by_op = collections.Counter(); samp_op = collections.Counter()
for r in body:
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDG", "LDS", "STS", "MUFU", "IMAD")) and "." in op else "")
    by_op[op] += int(r[ix["Instructions Executed"]]); samp_op[op] += int(r[ix["# Samples"]])
print("opcode            inst%   samples%")
for op, c in by_op.most_common(28):
    print(f"  {op:16s} {100 * c / tot_inst:6.2f}  {100 * samp_op[op] / max(tot_samp, 1):6.2f}")
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = {n: sum(int(r[ix[n]] or 0) for r in body) for n in stall_cols}
print("stall reasons (all samples):", {n[6:]: round(100 * v / max(tot_samp, 1), 1) for n, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
if "--list" in sys.argv:
    for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:40]:
        top = max(stall_cols, key=lambda n: int(r[ix[n]] or 0))
        print(f"{r[ix['Address']][-5:]} {int(r[ix['# Samples']]):7d} {int(r[ix['Instructions Executed']]):9d} {top[6:]:14s} {r[ix['Source']].strip()[:90]}")
