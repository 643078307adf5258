#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS view: `ncu -i rep --page source --csv --print-source sass`) by CUDA source
function and line, using the line table of the cubin (`nvdisasm -gi`).  Works for inlined code, where the symbol table
(ncu_by_function.py) sees only the kernel.

Usage: ncu_by_line.py <src.csv> <object-or-so> <kernel-substring> <n_reactors> <mean_steps> [top_lines]
"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

src, lib, kern, nre, nst = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), float(sys.argv[5])
top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubs = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]

# offset (within the kernel's .text section) -> (file, line) of the innermost frame
line_of = {}
for cub in cubs:
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cub], capture_output=True, text=True).stdout
    inside, cur, chain, fresh = False, ("?", 0), [], False
    for l in dis.splitlines():
        if l.startswith("//--------------------- .text."):
            inside = kern in l
            cur, chain, fresh = ("?", 0), [], False
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            # annotations between two instructions list the inline chain, innermost frame first
            if not fresh:
                chain, fresh = [], True
            chain.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", l)
        if m:
            if fresh:
                own = [c for c in chain if "/csrc/" in c[0]]
                f, ln = (own or chain)[0]
                cur = (os.path.basename(f), ln)
                fresh = False
            line_of[int(m.group(1), 16)] = cur
    if line_of:
        break

csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "reactionmechanismsimulator.jl_b200", "csrc")
fn_cache = {}
def func_of(f, ln):
    if f not in fn_cache:
        names = []
        try:
            lines = open(os.path.join(csrc, f)).read().splitlines()
        except OSError:
            lines = []
        for i, l in enumerate(lines, 1):
            s = l.strip()
            if s.startswith("//") or ("__device__" not in s and "__global__" not in s):
                continue
            # the name is the last identifier before the first '(' on this or one of the next two lines
            blob = " ".join(lines[i - 1:i + 2])
            blob = blob[blob.find("__device__") if "__device__" in blob else blob.find("__global__"):]
            blob = re.sub(r"__launch_bounds__\s*\([^)]*\)", "", blob)
            m = re.search(r"([A-Za-z_][A-Za-z0-9_:<>]*)\s*\(", blob)
            names.append((i, m.group(1) if m else "?"))
        fn_cache[f] = names
    name = "(top)"
    for i, n in fn_cache[f]:
        if i <= ln:
            name = n
        else:
            break
    return name

rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ie, ist = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ini = hdr.index("stall_no_inst") if "stall_no_inst" in hdr else None
base = int(rows[2][ia], 16)
by_fn, by_line = defaultdict(lambda: [0, 0, 0, 0]), defaultdict(lambda: [0, 0, 0])
for r in rows[2:]:
    off = int(r[ia], 16) - base
    f, ln = line_of.get(off, ("?", 0))
    fn = f + ":" + func_of(f, ln)
    ex, sm = int(r[ie]), int(r[ist] or 0)
    ni = int(r[ini] or 0) if ini is not None else 0
    a = by_fn[fn]; a[0] += ex; a[1] += sm; a[2] += 1; a[3] += ni
    b = by_line[(f, ln)]; b[0] += ex; b[1] += sm; b[2] += 1
tot = sum(v[0] for v in by_fn.values()); ts = sum(v[1] for v in by_fn.values()) or 1
print(f"total warp-instructions {tot:.4g}  ({tot / nre:.4g} per reactor, {tot / nre / nst:.0f} per step); static SASS {len(rows) - 2}")
print("   %inst  %stall  %no_inst  inst/step  static  function")
for n, v in sorted(by_fn.items(), key=lambda x: -x[1][0]):
    print(f"{v[0] / tot * 100:7.2f} {v[1] / ts * 100:7.2f} {v[3] / ts * 100:8.2f} {v[0] / nre / nst:9.0f} {v[2]:7d}  {n}")
print(f"\ntop {top} source lines by executed instructions")
for (f, ln), v in sorted(by_line.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{v[0] / tot * 100:6.2f}% inst {v[1] / ts * 100:6.2f}% stall {v[0] / nre / nst:7.0f}/step static {v[2]:5d}  {f}:{ln}")
