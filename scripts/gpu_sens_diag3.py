import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
import rmsb200 as rb  # noqa: E402
np.set_printoptions(linewidth=200)
gdom, y0, p = cases.product_domain("superminimal_V")
print(gdom.phase.names)
for tend in (0.03, 0.0384, 0.03845, 0.0385):
    for flags in (0, 1):
        sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p).ode, rb.RMSB200(max_steps=300000, flags=flags), abstol=1e-20, reltol=1e-6)
        st = sol.stats[0]
        print(f"flags {flags} tend {tend}: status {sol.status[0]} steps {st['n_steps']} rhs {st['n_rhs']} jac {st['n_jac']} nfail {st['n_newton_fail']} efail {st['n_err_fail']}")
        print("   yend", sol.yend[0])
J = gdom.jacobiany(sol.yend, p)[0]
print("J rows of species 0..3 max abs:", np.abs(J[:4]).max(axis=1))
print("f:", gdom.dydt(sol.yend, p)[0])
