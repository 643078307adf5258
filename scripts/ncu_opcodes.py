#!/usr/bin/env python
"""Executed-instruction mix per opcode of one kernel in an .ncu-rep (source page).  Usage: ncu_opcodes.py report.ncu-rep kernel_name [n_warps]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
nw = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "Instructions Executed" in r)
iS, iE, iSt = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
op, st, tot = collections.Counter(), collections.Counter(), 0
for r in rows:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    s = r[iS].strip().split()
    if not s:
        continue
    o = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    n = int(r[iE]); op[o] += n; st[o] += int(r[iSt]); tot += n
stot = sum(st.values()) or 1
print(f"{kern}: {tot} warp instructions ({tot / nw:.1f} per warp)")
for k, v in op.most_common(24):
    print(f"  {k:10s} {v / nw:9.1f} per warp  {100 * v / tot:5.1f}% inst  {100 * st[k] / stot:5.1f}% stall samples")
