"""Exploratory GPU check of config C5 (not a pytest file): jacobianp vs the oracle evaluated live, forward sensitivities
vs central finite differences of the GPU state solution, and sensitivity-ensemble throughput.
Run on the GPU box:  python scripts/gpu_sens_check.py"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))

import cases  # noqa: E402
import rmsb200 as rb  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402

for case in cases.CASES:
    odom = cases.oracle_domain(case)
    gdom, y0, p = cases.product_domain(case)
    ys = cases.states(case, odom.y0, odom.phase.Ns, odom.kind)[:2]
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.constantspeciesinds = gdom.constantspeciesinds
    ens.p, ens._p_default = gdom.p, gdom._p_default
    t0 = time.time()
    Jp = ens.jacobianp(ys, p)
    e = max(cases.err_jac(Jp[k], ro.jacobianp(odom, ys[k], 0.0, odom.p)) for k in range(len(ys)))
    print(f"jacp {case:16s} err {e:.2e}  shape {Jp.shape}", flush=True)
    if e > 1e-12:
        ref = ro.jacobianp(odom, ys[0], 0.0, odom.p)
        col = np.max(np.abs(ref), axis=0, keepdims=True)
        rel = np.abs(Jp[0] - ref) / (np.abs(ref) + 1e-3 * col + 1e-300)
        for idx in np.argsort(rel.ravel())[-6:]:
            i, j = np.unravel_index(idx, rel.shape)
            print(f"     row {i} col {j} (Ns={odom.phase.Ns}) gpu {Jp[0][i, j]:.15e} ref {ref[i, j]:.15e} colmax {col[0, j]:.3e}")

for case, tend in (("superminimal_TP", 150.11094), ("superminimal_V", 0.05), ("ethane_TP", 1e-3)):
    gdom, y0, p = cases.product_domain(case)
    t0 = time.time()
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p, forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-8)
    dt = time.time() - t0
    print(case, sol.retcode, "steps", sol.stats["n_steps"][0], "jac", sol.stats["n_jac"][0], f"{dt:.3f}s", flush=True)
    S = sol.sens_end[0]
    Ns = gdom.phase.Ns
    norms = np.abs(S).max(axis=1)
    for j in (int(np.argmax(norms[:Ns])), Ns + int(np.argmax(norms[Ns:])), Ns + int(np.argsort(norms[Ns:])[-2])):
        fd = []
        pj = np.asarray(p, dtype=float)[j]
        h = 1e-4 * abs(pj) if pj != 0 else 1e-4 * 8.314 * 1000
        for sgn in (1.0, -1.0):
            q = np.array(p, dtype=float)
            q[j] += sgn * h
            fd.append(rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=q).ode, abstol=1e-22, reltol=1e-10).yend[0])
        fd = (fd[0] - fd[1]) / (2 * h)
        print(f"   param {j}: |S|max {np.abs(S[j]).max():.3e}  FD rel err {np.abs(S[j] - fd).max() / np.abs(fd).max():.2e}", flush=True)

# throughput: C5 ensemble
gdom, y0, p = cases.product_domain("superminimal_TP")
ph = gdom.phase
for n in (2000, 20000):
    rng = np.random.default_rng(1)
    T = rng.uniform(950, 1250, n); P = rng.uniform(0.5e5, 5e5, n)
    ns = np.zeros((n, ph.Ns)); ns[:, ph.names.index("H2")] = rng.uniform(0.3, 0.7, n); ns[:, ph.names.index("O2")] = rng.uniform(0.2, 0.4, n)
    ens = rb.ConstantTPEnsemble(ph, T, P, ns)
    for rep in range(2):
        t0 = time.time()
        sol = rb.solve(rb.Reactor(ens, None, (0.0, 1.0), forwardsensitivities=True).ode, abstol=1e-21, reltol=1e-7)
        dt = time.time() - t0
    t0 = time.time()
    sol0 = rb.solve(rb.Reactor(ens, None, (0.0, 1.0)).ode, abstol=1e-21, reltol=1e-7)
    dt0 = time.time() - t0
    print(f"C5 ensemble n={n}: sens {n / dt:.0f} solves/s (steps mean {sol.stats['n_steps'].mean():.0f}), plain {n / dt0:.0f} solves/s, ok {np.mean(sol.status == 0):.3f}", flush=True)
