import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
import rmsb200 as rb  # noqa: E402

gdom, y0, p = cases.product_domain("superminimal_V")
for fs in (False, True):
    for rtol in (1e-6, 1e-7, 1e-8, 1e-12):
        for tend in (0.03, 0.0385, 0.05):
            sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p, forwardsensitivities=fs).ode, rb.RMSB200(max_steps=300000), abstol=1e-20, reltol=rtol)
            st = sol.stats[0]
            print(f"sens={fs} rtol {rtol:.0e} tend {tend}: status {sol.status[0]} steps {st['n_steps']} rhs {st['n_rhs']} jac {st['n_jac']} lu {st['n_lu']} nfail {st['n_newton_fail']} efail {st['n_err_fail']} order {st['last_order']} h {st['last_h']:.3e} T {sol.yend[0][gdom.phase.Ns]:.3f}", flush=True)
