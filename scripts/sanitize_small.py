"""Small instances of every kernel, for compute-sanitizer (memcheck / racecheck) on the GPU box:
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
import synth  # noqa: E402
import rmsb200 as rb  # noqa: E402

for case in ("superminimal_V", "ethane_TP", "liquid_TV", "surface_TAPhi"):
    gdom, y0, p = cases.product_domain(case)
    ys = np.repeat(gdom.y0, 5, axis=0)
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, 5, axis=0))
    ens.constantspeciesinds = gdom.constantspeciesinds
    ens.p, ens._p_default = gdom.p, gdom._p_default
    ens.dydt(ys, p); ens.jacobiany(ys, p); ens.jacobianp(ys, p); ens.rates(ys, p)
    tend = {"superminimal_V": 0.05, "ethane_TP": 1e-4, "liquid_TV": 10.0, "surface_TAPhi": 1e-6}[case]
    sol = rb.solve(rb.Reactor(ens, None, (0.0, tend), p=p).ode, abstol=1e-16, reltol=1e-6, saveat=[0.5 * tend, tend])
    print(case, "solve", sol.retcode[0], flush=True)
gdom, y0, p = cases.product_domain("superminimal_TP")
sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 5.0), p=p, forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-7, saveat=[1.0, 5.0])
print("sens", sol.retcode, flush=True)
sys.path.insert(0, str(ROOT / "tests"))
from test_gpu_coupled import build  # noqa: E402
react, y0, p, doms, inter = build()
sol = rb.solve(react.ode, abstol=1e-20, reltol=1e-6, saveat=[0.5e-5])
print("coupled", sol.retcode, flush=True)
path = "/tmp/synth160.rms"
synth.write_synthetic_mechanism(path, ns=160, seed=7)
d = rb.readinput(path)
ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase", diffusionlimited=False)
gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds={"T": 650.0, "V": 1.0, **synth.initial_moles(160)})
ys = np.repeat(gdom.y0, 3, axis=0)
ens = rb.Domain(gdom.kind, ph, ys, np.repeat(gdom.cond, 3, axis=0))
ens.dydt(ys); ens.jacobiany(ys)
sol = rb.solve(rb.Reactor(ens, None, (0.0, 1e-4)).ode, abstol=1e-12, reltol=1e-6, saveat=[1e-4])
print("large", sol.retcode, sol.stats["n_steps"], flush=True)
