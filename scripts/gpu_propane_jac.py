"""Jacobian of the propane ConstantV case at states along its ignition trajectory: GPU vs oracle (column-scaled and floor-based errors)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import cases
import rmsb200 as rb
from oracle import rms_oracle as ro
gdom, y0, p = cases.product_domain("propane_V")
odom = cases.oracle_domain("propane_V")
tq = np.array([1e-4, 5e-4, 1e-3, 2e-3])
sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 2.0e-3), p=p), abstol=1e-18, reltol=1e-8, saveat=tq)
ys = sol.u[0]
ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, len(ys), axis=0))
ens.p, ens._p_default = gdom.p, gdom._p_default
J = ens.jacobiany(ys, p)
for k, y in enumerate(ys):
    Jo = ro.jacobiany(odom, y, 0.0, odom.p)
    col = np.abs(Jo).max(axis=0, keepdims=True) + 1e-300
    print(f"t={tq[k]:g} T={y[-2]:.1f}  max |dJ|/colmax {np.max(np.abs(J[k]-Jo)/col):.3e}   max |dJ|/rowmax {np.max(np.abs(J[k]-Jo)/(np.abs(Jo).max(axis=1, keepdims=True)+1e-300)):.3e}  err_jac {cases.err_jac(J[k], Jo):.3e}")
