"""C3 at its stated size: distance of the GPU trajectory from the committed SciPy-BDF golden and from its own tighter runs
(how much of the distance is the golden's own global error).  Usage: python scripts/gpu_c3_golden_check.py"""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import rmsb200 as rb  # noqa: E402
import synth  # noqa: E402

with tempfile.TemporaryDirectory() as td:
    path = td + "/c3.rms"
    synth.write_octane_families(path, families=36)
    d = rb.readinput(path)
ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase")
ic = {"T": 450.0, "V": 1.0, **synth.octane_family_moles(36)}
gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds=ic, constantspecies=["oxygen"])
gold = np.load(ROOT / "tests" / "golden" / "c3_octane_families.npz")
tq, yq = gold["tq"], gold["yq"]
sols = {}
for rt in (1e-8, 1e-9, 1e-10, 1e-11, 1e-12):
    s = rb.solve(rb.Reactor(gdom, y0, (0.0, float(tq[-1])), p=p).ode, abstol=1e-18, reltol=rt, saveat=tq)
    sols[rt] = s.u[0]
    print(f"rtol {rt:g}: status {s.retcode} steps {int(s.stats['n_steps'][0])}", flush=True)


def dist(a, b):
    out = []
    for k in range(len(tq)):
        big = np.abs(b[k]) > 1e-9 * np.abs(b[k]).max()
        out.append(np.max(np.abs(a[k] - b[k])[big] / np.abs(b[k])[big]))
    return np.array(out)


for rt, u in sols.items():
    print(f"rtol {rt:g}: vs golden {dist(u, yq).max():.3e} (per save point {np.array2string(dist(u, yq), precision=2)})  vs GPU rtol 1e-12 {dist(u, sols[1e-12]).max():.3e}")
print(f"golden vs GPU rtol 1e-12: {dist(yq, sols[1e-12]).max():.3e}")
