"""One C3_35 ensemble solve (liquid octane, 35 species / 380 reactions, ConstantTV, diffusion limited: solve_kernel<2>) for ncu.
Usage: python scripts/prof_c3.py [n]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
import rmsb200 as rb  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4144
rng = np.random.default_rng(12)
gdom, _, _ = cases.product_domain("liquid_TV")
T = rng.uniform(420, 480, n)
y0s = np.repeat(gdom.y0, n, axis=0) * rng.uniform(0.8, 1.2, (n, 1))
ens = rb.Domain(gdom.kind, gdom.phase, y0s, np.stack([T, np.ones(n)], axis=1), constantspecies=["oxygen"])
for _ in range(2):
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 1.4e5)).ode, abstol=1e-16, reltol=1e-6)
print("ok", np.mean(sol.status == 0), sol.stats["n_steps"].mean(), "kernel ms", sol.kernel_ms, "solves/s", n / sol.kernel_ms * 1e3)
