"""ConstantV sensitivity diagnostic: tolerance scan + time history of S_T for one rate parameter."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
import rmsb200 as rb  # noqa: E402

gdom, y0, p = cases.product_domain("superminimal_V")
Ns = gdom.phase.Ns
tq = np.array([0.01, 0.03, 0.037, 0.038, 0.0383, 0.03838, 0.0384, 0.0385, 0.039, 0.04, 0.045, 0.05])
j = Ns + 4
for flags in (0, 3):
    for rtol in (1e-6, 1e-8, 1e-10, 1e-12):
        sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.05), p=p, forwardsensitivities=True).ode, rb.RMSB200(flags=flags), abstol=1e-20, reltol=rtol, saveat=tq)
        print(f"flags {flags} rtol {rtol:.0e} steps {sol.stats['n_steps'][0]} S_T[j={j}](t):", " ".join(f"{v:.4e}" for v in sol.sens[0, :, j, Ns]), flush=True)
# FD at every tq
fd = []
for sgn in (1.0, -1.0):
    q = np.array(p, dtype=float); q[j] += sgn * 1e-5
    fd.append(rb.solve(rb.Reactor(gdom, y0, (0.0, 0.05), p=q).ode, abstol=1e-22, reltol=1e-11, saveat=tq).u[0])
fd = (fd[0] - fd[1]) / 2e-5
print("FD S_T:", " ".join(f"{v:.4e}" for v in fd[:, Ns]))
sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.05), p=p, forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-10, saveat=tq)
print("all comps at t_end  S:", sol.sens[0, -1, j])
print("all comps at t_end FD:", fd[-1])
print("max order test:")
for mo in (1, 2, 3):
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.05), p=p, forwardsensitivities=True).ode, rb.RMSB200(max_order=mo), abstol=1e-20, reltol=1e-8, saveat=tq)
    print(f"max_order {mo} steps {sol.stats['n_steps'][0]} S_T:", " ".join(f"{v:.4e}" for v in sol.sens[0, :, j, Ns]), flush=True)
