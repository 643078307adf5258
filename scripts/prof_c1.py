"""One C1 ensemble solve (H2/O2 ConstantTP: solve_kernel<1>) for ncu.  Usage: python scripts/prof_c1.py [n]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
import rmsb200 as rb  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(11)
gdom, y0, p = cases.product_domain("superminimal_TP")
ph = gdom.phase
T = rng.uniform(950, 1250, n); P = rng.uniform(0.5e5, 5e5, n)
ns = np.zeros((n, ph.Ns)); ns[:, ph.names.index("H2")] = rng.uniform(0.3, 0.7, n); ns[:, ph.names.index("O2")] = rng.uniform(0.2, 0.4, n)
ens = rb.ConstantTPEnsemble(ph, T, P, ns)
for _ in range(2):
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 1.0)).ode, abstol=1e-16, reltol=1e-6)
print("ok", np.mean(sol.status == 0), sol.stats["n_steps"].mean(), "kernel ms", sol.kernel_ms, "solves/s", n / sol.kernel_ms * 1e3)
