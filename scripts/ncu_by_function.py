#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (per-SASS-instruction counts) by device function, using the cubin's
symbol table.  Usage: ncu_by_function.py <src.csv> <lib.so> <kernel-substring> <n_reactors> <mean_steps>"""
import csv, subprocess, sys, tempfile, os
src, lib, kern, nre, nst = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), float(sys.argv[5])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
out = subprocess.run(["readelf", "-sW", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
syms = []
for l in out.splitlines():
    p = l.split()
    if len(p) >= 8 and p[3] == "FUNC" and kern in p[-1]:
        size = int(p[2], 16) if p[2].startswith("0x") else int(p[2])
        name = p[-1].split("$")[-1] if "$" in p[-1] else "MAIN"
        syms.append((int(p[1], 16), size, name))
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ie, ist = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
base = int(rows[2][ia], 16)
main_start = [v for v, s, n in syms if n == "MAIN"][0]
agg, samp = {}, {}
for r in rows[2:]:
    off = int(r[ia], 16) - base + main_start
    name = "MAIN"
    for v, s, n in syms:
        if n != "MAIN" and v <= off < v + s:
            name = n
            break
    agg[name] = agg.get(name, 0) + int(r[ie])
    samp[name] = samp.get(name, 0) + int(r[ist] or 0)
tot, ts = sum(agg.values()), sum(samp.values())
print(f"total warp-instructions {tot:.4g}  ({tot / nre:.4g} per reactor, {tot / nre / nst:.0f} per step)")
for n, v in sorted(agg.items(), key=lambda x: -x[1]):
    print(f"{v / tot * 100:6.2f}% inst {samp[n] / ts * 100:6.2f}% stall-samples {v / nre / nst:8.0f} inst/step  {n[:80]}")
