"""Exploratory check of the large-mechanism path (config C3) on a synthetic mechanism (not a pytest file).
Usage: python scripts/gpu_large_check.py [ns]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
import synth  # noqa: E402
import rmsb200 as rb  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 300
path = f"/tmp/synth{ns}.rms"
synth.write_synthetic_mechanism(path, ns=ns, seed=7)
t0 = time.time()
d = rb.readinput(path)
ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase", diffusionlimited=False)
ic = {"T": 650.0, "V": 1.0, **synth.initial_moles(ns)}
gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds=ic)
print(f"product set-up {time.time() - t0:.1f}s  Ns {ph.Ns} Nr {ph.Nr}", flush=True)
from rmsb200 import _lib as L  # noqa: E402
info = np.zeros(8, dtype=np.int64)
L.check(L.lib().rmsb_large_info(ph.handle, L.ptr(info)))
print("large info: n %d nnzLU %d lvlL %d lvlU %d maxrow %d factor_fma %d core_start %d scratch_doubles %d" % tuple(info), flush=True)


def profile(tag):
    cyc = np.zeros(8)
    L.check(L.lib().rmsb_large_profile(ph.handle, 1, L.ptr(cyc)))
    tot = max(cyc[4], 1.0)
    print(f"   profile {tag}: rhs {cyc[0] / tot:.2f} jac {cyc[1] / tot:.2f} factor {cyc[2] / tot:.2f} solve {cyc[3] / tot:.2f} | factor parts: sparse rows {cyc[5] / tot:.2f} +core rows {(cyc[6] - cyc[5]) / tot:.2f} dense core {cyc[7] / tot:.2f} (of {tot:.3e} block-cycles)", flush=True)
od = ro.readinput(path)
oph = ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0], diffusionlimited=False)
odom = ro.ConstantTVDomain(oph, ic)
rng = np.random.default_rng(1)
ys = np.stack([odom.y0, np.abs(odom.y0 + 10.0 ** rng.uniform(-3, 1, ns))])
ens2 = rb.Domain(gdom.kind, ph, ys, np.repeat(gdom.cond, 2, axis=0))
f = ens2.dydt(ys)
for k in range(2):
    fo = ro.dydtreactor(odom, ys[k], 0.0, None)
    print("dydt err", cases.err_dydt(f[k], fo), flush=True)
np.testing.assert_allclose(gdom.kfs[0], odom.kfs, rtol=1e-12)
np.testing.assert_allclose(gdom.krevs[0], odom.krevs, rtol=1e-12)
if ns <= 400:
    J = ens2.jacobiany(ys)
    for k in range(2):
        print("jac err", cases.err_jac(J[k], ro.jacobiany(odom, ys[k], 0.0, None)), flush=True)
cm = co.CpuMech(oph)
for tend, rtol in ((1.0, 1e-6),) if ns > 400 else ((1e-3, 1e-6), (1.0, 1e-6), (1e-3, 1e-9)):
    t0 = time.time()
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p).ode, abstol=1e-12, reltol=rtol)
    tg = time.time() - t0
    t0 = time.time()
    ref = cm.solve("ConstantTV", odom.y0[None], np.array([[650.0, 1.0]]), tend, rtol, 1e-12, nthreads=1)
    tc = time.time() - t0
    big = np.abs(ref["yend"][0]) > 1e-8 * np.abs(ref["yend"][0]).max()
    err = np.max(np.abs(sol.yend[0] - ref["yend"][0])[big] / np.abs(ref["yend"][0])[big])
    profile("single")
    print(f"tend {tend} rtol {rtol}: gpu {sol.retcode} {sol.stats[0]} {tg:.2f}s | cpu {ref['stats'][0]} {tc:.2f}s | max rel diff {err:.2e}", flush=True)
# throughput: ensemble of conditions
n = int(sys.argv[2]) if len(sys.argv) > 2 else 296
T = np.random.default_rng(2).uniform(600.0, 700.0, n)
y0s = np.repeat(gdom.y0, n, axis=0) * np.random.default_rng(3).uniform(0.8, 1.2, (n, 1))
ens = rb.Domain(gdom.kind, ph, y0s, np.stack([T, np.ones(n)], axis=1))
for rep in range(2):
    t0 = time.time()
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 1.0)).ode, abstol=1e-12, reltol=1e-6)
    dt = time.time() - t0
profile("ensemble")
print("status counts:", np.unique(sol.status, return_counts=True), "failed stats:", sol.stats[sol.status != 0][:3])
print(f"ensemble n={n}: {n / dt:.1f} solves/s, ok {np.mean(sol.status == 0):.3f}, mean steps {sol.stats['n_steps'].mean():.0f} lu {sol.stats['n_lu'].mean():.0f} rhs {sol.stats['n_rhs'].mean():.0f}")
