"""Integrator-flag sweep on the coupled gas + catalyst ensemble (C4) and the 35-species liquid ensemble (C3_35): do the Newton
tolerance / test variants tuned on C2 suit the warp-per-reactor path?  flags: bit0 scipy-style Newton test, bit1 no step hysteresis,
bits 8..15 = v: Newton tolerance 0.1 / 2^(v-1) (0 = default 0.0125).  Usage: python scripts/gpu_flags_c4.py [n_c4] [n_c3]"""
import sys
import time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
import rmsb200 as rb  # noqa: E402
from test_gpu_coupled import build  # noqa: E402

n4 = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
n3 = int(sys.argv[2]) if len(sys.argv) > 2 else 4736
rng = np.random.default_rng(14)
react, y0, p, doms, inter = build()
cph = inter.phase
N1, N2 = doms[0].phase.Ns, doms[1].phase.Ns
T = rng.uniform(700, 1000, n4); x = rng.uniform(0.05, 0.2, n4)
nsg = np.zeros((n4, N1)); nss = np.zeros((n4, N2))
gn, sn = doms[0].phase.names, doms[1].phase.names
nsg[:, gn.index("O2")], nsg[:, gn.index("N2")], nsg[:, gn.index("CH4")] = 0.2, 0.7, x
Pp = np.full(n4, 1.0e5)
V = nsg.sum(axis=1) * 8.314 * T / Pp
A = 1.0e5 * V
nss[:, sn.index("vacantX")] = 2.486e-5 * A
ens4 = rb.CoupledEnsemble(cph, T, Pp, A, nsg, nss)
rng = np.random.default_rng(12)
gdom, _, _ = cases.product_domain("liquid_TV")
T3 = rng.uniform(420, 480, n3)
y0s = np.repeat(gdom.y0, n3, axis=0) * rng.uniform(0.8, 1.2, (n3, 1))
ens3 = rb.Domain(gdom.kind, gdom.phase, y0s, np.stack([T3, np.ones(n3)], axis=1), constantspecies=["oxygen"])
ref = {}
for name, mk, tol in (("C4", lambda: rb.Reactor(ens4, None, (0.0, 0.1)), dict(abstol=1e-20, reltol=1e-6)),
                      ("C3_35", lambda: rb.Reactor(ens3, None, (0.0, 1.4e5)).ode, dict(abstol=1e-16, reltol=1e-6))):
    tight = rb.solve(mk(), rb.RMSB200(), abstol=tol["abstol"], reltol=1e-10)
    for flags in (0, 1, 2, 0x100, 0x200, 0x300, 0x500, 0x600, 0x101):
        rb.solve(mk(), rb.RMSB200(flags=flags), **tol)
        sol = rb.solve(mk(), rb.RMSB200(flags=flags), **tol)
        st = sol.stats
        big = np.abs(tight.yend) > 1e-6 * np.abs(tight.yend).max(axis=1, keepdims=True)
        err = np.max(np.abs(sol.yend - tight.yend)[big] / np.abs(tight.yend)[big])
        print(f"{name} flags={flags:#06x}: {len(sol.status) / sol.kernel_ms * 1e3:9.0f} solves/s  ok {np.mean(sol.status == 0):.3f}  steps {st['n_steps'].mean():7.1f} rhs {st['n_rhs'].mean():7.1f} "
              f"jac {st['n_jac'].mean():6.1f} lu {st['n_lu'].mean():6.1f} nfail {st['n_newton_fail'].mean():6.1f} efail {st['n_err_fail'].mean():5.1f}  max err vs rtol 1e-10: {err:.2e}", flush=True)
