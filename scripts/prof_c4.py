"""One coupled gas + Pt-surface ensemble solve (C4: Ny 42, two rows per lane) for ncu.  Usage: python scripts/prof_c4.py [n]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import rmsb200 as rb  # noqa: E402
from test_gpu_coupled import build  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
react, y0, p, doms, inter = build()
cph = inter.phase
N1, N2 = doms[0].phase.Ns, doms[1].phase.Ns
rng = np.random.default_rng(3)
T = rng.uniform(700, 1000, n); x = rng.uniform(0.05, 0.2, n)
nsg = np.zeros((n, N1)); nss = np.zeros((n, N2))
gn, sn = doms[0].phase.names, doms[1].phase.names
nsg[:, gn.index("O2")], nsg[:, gn.index("N2")], nsg[:, gn.index("CH4")] = 0.2, 0.7, x
Pp = np.full(n, 1.0e5)
V = nsg.sum(axis=1) * 8.314 * T / Pp
A = 1.0e5 * V
nss[:, sn.index("vacantX")] = 2.486e-5 * A
ens = rb.CoupledEnsemble(cph, T, Pp, A, nsg, nss)
for _ in range(2):
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-20, reltol=1e-6)
print("ok", np.mean(sol.status == 0), sol.stats["n_steps"].mean())
