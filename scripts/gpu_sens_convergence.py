"""C5 forward sensitivities: distance of the GPU result from the oracle's augmented-system golden (SciPy BDF, rtol 1e-9) as the GPU
tolerance tightens (column-scaled max error at every saved time).  Usage: python scripts/gpu_sens_convergence.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import cases
import rmsb200 as rb
G = np.load(ROOT / "tests" / "golden" / "sens_vectors.npz")


def sens_err(S, Sref):
    sc = np.maximum(np.max(np.abs(Sref), axis=-1, keepdims=True), 1e-25 * np.max(np.abs(Sref))) + 1e-300
    return float(np.max(np.abs(S - Sref) / sc))


gdom, y0, p = cases.product_domain("superminimal_TP")
tq, Sq = G["superminimal_TP/tq"], G["superminimal_TP/Sq"]
for rt in (1e-6, 1e-7, 1e-8, 1e-9, 1e-10):
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 150.11094), p=p, forwardsensitivities=True).ode, abstol=1e-22, reltol=rt, saveat=tq)
    print(f"rtol {rt:g}: steps {int(sol.stats['n_steps'][0])}  max column-scaled error vs golden {max(sens_err(sol.sens[0, k], Sq[k]) for k in range(len(tq))):.2e}")
