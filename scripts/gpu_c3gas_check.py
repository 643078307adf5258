"""Exploratory check of the large-mechanism path on VARIABLE-k gas domains with falloff / third-body kinetics: the octane-family
mechanism read as an IdealGas (tests/synth.py::write_octane_families(falloff=True)): dydt / Jacobian / kf / krev parity against the NumPy
oracle for ConstantV, ConstantP and ConstantTP, one ConstantV trajectory against the C++ CPU port (dense pivoted LU), a small sweep.
Usage (GPU box): python scripts/gpu_c3gas_check.py [n_conditions] [families]"""
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
from oracle import floors as fl  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402
from rmsb200 import _lib as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
K = int(sys.argv[2]) if len(sys.argv) > 2 else 36
path = f"/tmp/octgas{K}.rms"
info = synth.write_octane_families(path, families=K, falloff=True)
d = rb.readinput(path)
ph = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
od = ro.readinput(path)
oph = ro.IdealGas(od["phase"]["Species"], od["phase"]["Reactions"])
Ns = ph.Ns
fuel = {k: v / 6154.0 for k, v in synth.octane_family_moles(K).items() if k != "oxygen"}
ic = {"T": 1000.0, "P": 20e5, **fuel, "oxygen": 6.25, "N2": 23.5}
print(f"Ns {ph.Ns} Nr {ph.Nr} efficiency reactions {len(ph.efficiencyinds)}", flush=True)
rng = np.random.default_rng(1)
for name, gctor, octor in (("ConstantV", rb.ConstantVDomain, ro.ConstantVDomain), ("ConstantP", rb.ConstantPDomain, ro.ConstantPDomain),
                           ("ConstantTP", rb.ConstantTPDomain, ro.ConstantTPDomain)):
    gdom, y0, p = gctor(phase=ph, initialconds=ic)
    odom = octor(oph, ic)
    ys = np.stack([odom.y0, np.abs(odom.y0 * rng.uniform(0.5, 1.5, len(odom.y0)) + np.r_[1e-6 * np.ones(Ns), np.zeros(len(odom.y0) - Ns)])])
    ens = rb.Domain(gdom.kind, ph, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.p, ens._p_default = gdom.p, gdom._p_default
    t0 = time.time()
    f, J = ens.dydt(ys, p), ens.jacobiany(ys, p)
    kf, kr = ens.rates(ys, p)
    for k in range(len(ys)):
        po = odom.p
        fo, M = fl.dydt_terms(odom, ys[k], 0.0, po)
        rep = fl.report(f[k], fo, fl.EPS * M, 1e-12, slack=4.0)
        Jo = ro.jacobiany(odom, ys[k], 0.0, po)
        th = odom.calcthermo(ys[k], 0.0, po)
        print(f"{name} state {k}: dydt rel above floor {rep['max_rel_above']:.2e} excess {rep['worst_excess']:.2e} | J mixed err {cases.err_jac(J[k], Jo):.2e} "
              f"zero-pattern ok {bool(np.all(J[k][Jo == 0.0] == 0.0))} | kf {np.max(np.abs(kf[k] - th.kfs) / np.abs(th.kfs)):.2e} krev {np.max(np.abs(kr[k] - th.krevs) / (np.abs(th.krevs) + 1e-300)):.2e}",
              flush=True)
        bad = np.argwhere(~(np.abs(J[k] - Jo) <= 1e-12 * (np.abs(Jo) + 1e-3 * np.abs(Jo).max(axis=0, keepdims=True))))
        if len(bad):
            print("   entries beyond 1e-12 mixed:", len(bad), "first (row, col, gpu, oracle):", [(int(a), int(b), J[k][a, b], Jo[a, b]) for a, b in bad[:6]], "Ns", Ns)
        zb = np.argwhere((Jo == 0.0) & (J[k] != 0.0))
        if len(zb):
            print("   nonzero where the oracle is exactly zero:", len(zb), [(int(a), int(b), J[k][a, b], np.abs(Jo[:, b]).max()) for a, b in zb[:6]])
        if k == 1:
            Jr, Jf = fl.jac_floor(lambda yy, pp: ro.jacobiany(odom, yy, 0.0, pp), odom, ys[k], po, nsample=3)
            print("   floor-based:", {a: b for a, b in fl.report(J[k], Jr, Jf, 1e-12).items()})
# ConstantV trajectory vs the CPU port
gdom, y0, p = rb.ConstantVDomain(phase=ph, initialconds=ic)
odom = ro.ConstantVDomain(oph, ic)
tend = 2e-3
tq = np.array([0.25 * tend, tend])
t0 = time.time()
sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p).ode, abstol=1e-16, reltol=1e-8, saveat=tq)
print("gpu solve", time.time() - t0, sol.retcode, sol.stats[0], "T_end", sol.yend[0, Ns], "idt", sol.idt[0], flush=True)
t0 = time.time()
ref = co.CpuMech(oph).solve("ConstantV", odom.y0[None], odom.cond[None], tend, 1e-8, 1e-16, tsave=tq, nthreads=1)
print("cpu port", time.time() - t0, ref["status"], ref["stats"][0], "T_end", ref["yend"][0, Ns], flush=True)
for k in range(len(tq)):
    yr = ref["ysave"][0, k]
    big = np.abs(yr) > 1e-9 * np.abs(yr[:Ns]).max()
    print(" t", tq[k], "max rel diff", np.max(np.abs(sol.u[0, k] - yr)[big] / np.abs(yr)[big]))
# sweep
T0 = rng.uniform(900.0, 1200.0, n); P0 = 10 ** rng.uniform(5.7, 6.7, n)
nsv = np.repeat(gdom.y0[:, :Ns], n, axis=0)
ens = rb.ConstantVEnsemble(ph, T0, P0, nsv / nsv.sum(axis=1, keepdims=True))
for rep_ in range(2):
    t0 = time.time()
    s = rb.solve(rb.Reactor(ens, None, (0.0, 1e-2)).ode, abstol=1e-16, reltol=1e-6)
    dt = time.time() - t0
print(f"ConstantV sweep n={n}: {n / dt:.1f} solves/s  status hist {np.bincount(s.status, minlength=4)}  mean steps {s.stats['n_steps'].mean():.0f} rhs {s.stats['n_rhs'].mean():.0f} "
      f"jac {s.stats['n_jac'].mean():.1f} lu {s.stats['n_lu'].mean():.0f}; ignited {np.isfinite(s.idt).sum()} T_end range {s.yend[:, Ns].min():.0f}-{s.yend[:, Ns].max():.0f}")
