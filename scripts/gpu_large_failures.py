"""Find failing reactors of a synthetic-mechanism ensemble on the large path and compare with the CPU port."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import synth  # noqa: E402
import rmsb200 as rb  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402

ns = int(sys.argv[1]); n = int(sys.argv[2])
path = f"/tmp/synth{ns}.rms"
synth.write_synthetic_mechanism(path, ns=ns, seed=7)
d = rb.readinput(path)
ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase", diffusionlimited=False)
ic = {"T": 650.0, "V": 1.0, **synth.initial_moles(ns)}
gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds=ic)
T = np.random.default_rng(2).uniform(600.0, 700.0, n)
y0s = np.repeat(gdom.y0, n, axis=0) * np.random.default_rng(3).uniform(0.8, 1.2, (n, 1))
ens = rb.Domain(gdom.kind, ph, y0s, np.stack([T, np.ones(n)], axis=1))
sol = rb.solve(rb.Reactor(ens, None, (0.0, 1.0)).ode, abstol=1e-12, reltol=1e-6)
bad = np.flatnonzero(sol.status != 0)
print("failed:", bad, "T:", T[bad], "steps", sol.stats["n_steps"][bad])
od = ro.readinput(path)
oph = ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0], diffusionlimited=False)
cm = co.CpuMech(oph)
for r in list(bad[:3]) + [0]:
    ref = cm.solve("ConstantTV", y0s[r:r + 1], np.array([[T[r], 1.0]]), 1.0, 1e-6, 1e-12, nthreads=1)
    print("r", r, "gpu status", sol.status[r], sol.stats[r], "| cpu status", ref["status"][0], ref["stats"][0], "min y gpu", sol.yend[r].min(), "cpu", ref["yend"][0].min())
    for atol in (1e-10, 1e-8):
        one = rb.Domain(gdom.kind, ph, y0s[r:r + 1], np.array([[T[r], 1.0]]))
        s1 = rb.solve(rb.Reactor(one, None, (0.0, 1.0)).ode, abstol=atol, reltol=1e-6)
        print("    gpu atol", atol, s1.status[0], s1.stats[0])
