"""Exploratory check of config C3 at its stated size (tests/synth.py::write_octane_families: 1015 species / 11.6 k reactions):
pattern analysis, dydt / Jacobian parity against the NumPy oracle, one trajectory against SciPy BDF with a sparse Jacobian, and
the throughput of a T sweep.  Usage (GPU box): python scripts/gpu_c3_check.py [n_conditions] [families]"""
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
from oracle import rms_oracle as ro  # noqa: E402
from rmsb200 import _lib as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
K = int(sys.argv[2]) if len(sys.argv) > 2 else 36
path = f"/tmp/oct{K}.rms"
info = synth.write_octane_families(path, families=K)
t0 = time.time()
d = rb.readinput(path)
ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase")
od = ro.readinput(path)
oph = ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0])
ic = {"T": 450.0, "V": 1.0, **synth.octane_family_moles(K)}
gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds=ic, constantspecies=["oxygen"])
odom = ro.ConstantTVDomain(oph, ic, constantspecies=("oxygen",))
li = np.zeros(8, dtype=np.int64)
L.check(L.lib().rmsb_large_info(ph.handle, L.ptr(li)))
print(f"Ns {ph.Ns} Nr {ph.Nr}; setup {time.time() - t0:.1f} s; large info n {li[0]} nnz {li[1]} levels {li[2]}/{li[3]} max_row {li[4]} fma {li[5]} core {li[6]} scratch {li[7]}", flush=True)
Ns = ph.Ns
rng = np.random.default_rng(1)
ys = np.stack([odom.y0, np.abs(odom.y0 + 10.0 ** rng.uniform(-3, 1, Ns)), np.abs(odom.y0 * rng.uniform(0.5, 1.5, Ns) + 1e-6)])
ens = rb.Domain(gdom.kind, ph, ys, np.repeat(gdom.cond, len(ys), axis=0), constantspecies=["oxygen"])
ens.p, ens._p_default = gdom.p, gdom._p_default
f, J = ens.dydt(ys, p), ens.jacobiany(ys, p)
for k in range(len(ys)):
    print("state", k, "dydt err", cases.err_dydt(f[k], ro.dydtreactor(odom, ys[k], 0.0, odom.p)), "jac err", cases.err_jac(J[k], ro.jacobiany(odom, ys[k], 0.0, odom.p)), flush=True)
print("kf rel", np.max(np.abs(gdom.kfs[0] - odom.kfs) / np.abs(odom.kfs)), "krev rel", np.max(np.abs(gdom.krevs[0] - odom.krevs) / (np.abs(odom.krevs) + 1e-300)))
# trajectory vs SciPy BDF (sparse Jacobian)
from scipy.integrate import solve_ivp  # noqa: E402
from scipy.sparse import csc_matrix  # noqa: E402
tq = np.array([1.0, 1e2, 1e4, 1.4e5])
t0 = time.time()
sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 1.4e5), p=p).ode, abstol=1e-16, reltol=1e-9, saveat=tq)
print("gpu single solve", time.time() - t0, sol.retcode, sol.stats[0], flush=True)
t0 = time.time()
ref = solve_ivp(lambda t, y: ro.dydtreactor(odom, y, t, odom.p), (0.0, 1.4e5), odom.y0, method="BDF", rtol=1e-9, atol=1e-16,
                jac=lambda t, y: csc_matrix(ro.jacobiany(odom, y, t, odom.p)), t_eval=tq)
print("scipy BDF", time.time() - t0, ref.status, ref.nfev, ref.njev, ref.nlu, flush=True)
for k in range(len(tq)):
    yr = ref.y[:, k]
    big = np.abs(yr) > 1e-9 * np.abs(yr).max()
    print(" t", tq[k], "max rel err", np.max(np.abs(sol.u[0, k] - yr)[big] / np.abs(yr)[big]), "x(octane fam 0)", yr[ph.names.index("F00_octane")] / yr.sum())
# sweep
T = rng.uniform(400.0, 500.0, n)
y0s = np.repeat(gdom.y0, n, axis=0) * rng.uniform(0.8, 1.2, (n, 1))
y0s[:, ph.names.index("oxygen")] = 4.953 * rng.uniform(0.5, 2.0, n)
ens = rb.Domain(gdom.kind, ph, y0s, np.stack([T, np.ones(n)], axis=1), constantspecies=["oxygen"])
for rep in range(2):
    t0 = time.time()
    s = rb.solve(rb.Reactor(ens, None, (0.0, 140000.01)).ode, abstol=1e-16, reltol=1e-6)
    dt = time.time() - t0
print(f"sweep n={n}: {n / dt:.1f} solves/s (kernel {s.kernel_ms:.1f} ms -> {n / s.kernel_ms * 1e3:.1f}/s)  status hist {np.bincount(s.status, minlength=4)}  mean steps {s.stats['n_steps'].mean():.0f} "
      f"rhs {s.stats['n_rhs'].mean():.0f} jac {s.stats['n_jac'].mean():.1f} lu {s.stats['n_lu'].mean():.0f} nfail {s.stats['n_newton_fail'].mean():.1f} efail {s.stats['n_err_fail'].mean():.1f}")
prof = np.zeros(8)
L.check(L.lib().rmsb_large_profile(ph.handle, 1, L.ptr(prof)))
print("cycle split rhs/jac/factor/solve/total:", prof[:5] / max(prof[4], 1))
