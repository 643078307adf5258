#!/usr/bin/env python
"""Per-instruction listing of one device function from an ncu source-page CSV, with executions per call.
Usage: ncu_func_listing.py <src.csv> <lib.so> <kernel-substring> <function-substring> <n_calls>"""
import csv, subprocess, sys, tempfile, os
src, lib, kern, fn, ncalls = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
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
ia, ie, isrc, ist = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
base = int(rows[2][ia], 16)
main_start = [v for v, s, n in syms if n == "MAIN"][0]
tgt = [(v, s) for v, s, n in syms if fn in n][0] if fn != "MAIN" else None
others = [(v, s) for v, s, n in syms if n != "MAIN"]
for r in rows[2:]:
    off = int(r[ia], 16) - base + main_start
    if tgt:
        inside = tgt[0] <= off < tgt[0] + tgt[1]
    else:
        inside = not any(v <= off < v + s for v, s in others)
    if inside:
        print(f"{off:6x} {int(r[ie]) / ncalls:7.2f} {int(r[ist] or 0):6d}  {r[isrc].strip()[:110]}")
