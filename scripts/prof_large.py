"""One large-path ensemble solve for ncu: C3 at its stated size (octane template x 36 families, 1015 species / 11 593 reactions,
ConstantTV, diffusion limited), one resident wave of reactors.  Usage: python scripts/prof_large.py [n]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import synth  # noqa: E402
import rmsb200 as rb  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
path = "/tmp/rmsb_octane_families_36.rms"
synth.write_octane_families(path, families=36)
d = rb.readinput(path)
ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase")
gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds={"T": 450.0, "V": 1.0, **synth.octane_family_moles(36)}, constantspecies=["oxygen"])
rng = np.random.default_rng(13)
T = rng.uniform(400.0, 500.0, n)
y0s = np.repeat(gdom.y0, n, axis=0) * rng.uniform(0.8, 1.2, (n, 1))
y0s[:, ph.names.index("oxygen")] = 4.953 * rng.uniform(0.5, 2.0, n)
ens = rb.Domain(gdom.kind, ph, y0s, np.stack([T, np.ones(n)], axis=1), constantspecies=["oxygen"])
for _ in range(2):
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 140000.01)).ode, abstol=1e-16, reltol=1e-6)
print("ok", np.mean(sol.status == 0), sol.stats["n_steps"].mean(), "kernel ms", sol.kernel_ms, "solves/s", n / sol.kernel_ms * 1e3)
