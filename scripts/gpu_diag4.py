import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
import rmsb200 as rb  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402
from test_gpu_parity import sweep  # noqa: E402
np.set_printoptions(linewidth=220)
gdom, _, _ = cases.product_domain("superminimal_V")
ig = gdom.phase
n = 20000
T0, P0, ns = sweep(ig, n)
ens = rb.ConstantVEnsemble(ig, T0, P0, ns)
sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-16, reltol=1e-6)
Ni = sol.yend[:, ig.Ns + 1] * ens.cond[:, 0] / (8.314 * sol.yend[:, ig.Ns])
err = np.abs(Ni / sol.yend[:, :ig.Ns].sum(axis=1) - 1)
print("frac > 1e-5:", np.mean(err > 1e-5), "max", err.max(), "steps mean", sol.stats["n_steps"].mean())
worst = np.argsort(err)[-3:]
odom = cases.oracle_domain("superminimal_V")
cm = co.CpuMech(odom.phase)
for r in worst:
    st = sol.stats[r]
    print(f"r={r} err {err[r]:.2e} T0 {T0[r]:.1f} P0 {P0[r]:.3e} N2 {ns[r, ig.spcdict['N2'] - 1]:.3f} steps {st['n_steps']} rhs {st['n_rhs']} jac {st['n_jac']} lu {st['n_lu']} nfail {st['n_newton_fail']} efail {st['n_err_fail']} Tend {sol.yend[r, ig.Ns]:.4f} idt {sol.idt[r]:.6e}")
    one = rb.Domain(ens.kind, ig, ens.y0[r:r + 1], ens.cond[r:r + 1]); one.p, one._p_default = ens.p, ens._p_default
    for rtol, flags in ((1e-6, 0), (1e-6, 4), (1e-6, 8), (1e-6, 12), (1e-6, 1), (1e-6, 2)):
        s1 = rb.solve(rb.Reactor(one, None, (0.0, 0.1)), rb.RMSB200(flags=flags), abstol=1e-16, reltol=rtol)
        N1 = s1.yend[0, ig.Ns + 1] * one.cond[0, 0] / (8.314 * s1.yend[0, ig.Ns])
        print(f"     rtol {rtol:.0e} flags {flags}: steps {s1.stats['n_steps'][0]} Tend {s1.yend[0, ig.Ns]:.4f} inv err {abs(N1 / s1.yend[0, :ig.Ns].sum() - 1):.2e} idt {s1.idt[0]:.6e}")
    ref = cm.solve("ConstantV", ens.y0[r:r + 1], ens.cond[r:r + 1], 0.1, 1e-6, 1e-16, nthreads=1)
    Nr_ = ref["yend"][0, ig.Ns + 1] * one.cond[0, 0] / (8.314 * ref["yend"][0, ig.Ns])
    print(f"     cpu rtol 1e-6: steps {ref['stats']['n_steps'][0]} Tend {ref['yend'][0, ig.Ns]:.4f} inv err {abs(Nr_ / ref['yend'][0, :ig.Ns].sum() - 1):.2e}")
# time history of the invariant for the worst
r = worst[-1]
one = rb.Domain(ens.kind, ig, ens.y0[r:r + 1], ens.cond[r:r + 1]); one.p, one._p_default = ens.p, ens._p_default
idt = sol.idt[r]
tq = np.concatenate([idt * np.array([0.5, 0.9, 0.99, 0.999, 1.0, 1.001, 1.01, 1.1, 2.0]), [0.1]])
tq = np.unique(np.minimum(tq, 0.1))
s1 = rb.solve(rb.Reactor(one, None, (0.0, 0.1)), abstol=1e-16, reltol=1e-6, saveat=tq)
for k, t in enumerate(tq):
    u = s1.u[0, k]
    print(f"   t {t:.6e} T {u[ig.Ns]:.3f} inv err {u[ig.Ns + 1] * one.cond[0, 0] / (8.314 * u[ig.Ns]) / u[:ig.Ns].sum() - 1:.3e}")
