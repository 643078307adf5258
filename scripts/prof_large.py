"""One large-path ensemble solve (synthetic 300 species, 592 reactors) for ncu.  Usage: python scripts/prof_large.py"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import synth  # noqa: E402
import rmsb200 as rb  # noqa: E402
ns, n = 300, 592
path = f"/tmp/synth{ns}.rms"
synth.write_synthetic_mechanism(path, ns=ns, seed=7)
d = rb.readinput(path)
ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase", diffusionlimited=False)
gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds={"T": 650.0, "V": 1.0, **synth.initial_moles(ns)})
rng = np.random.default_rng(11)
ens = rb.Domain(gdom.kind, ph, np.repeat(gdom.y0, n, axis=0) * rng.uniform(0.8, 1.2, (n, 1)), np.stack([rng.uniform(600, 700, n), np.ones(n)], axis=1))
for _ in range(2):
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 1.0)).ode, abstol=1e-12, reltol=1e-6)
print("ok", np.mean(sol.status == 0), sol.stats["n_steps"].mean())
