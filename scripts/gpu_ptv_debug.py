import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import cases
import rmsb200 as rb
IC_TV = {"ts": [0.0, 600.0, 1200.0], "T": [450.0, 490.0, 500.0], "V": 1.0, "octane": 6154.0, "oxygen": 4.953}
d = rb.readinput(cases.fixture("liquid_phase"))
liq = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase")
gdom, y0, p = rb.ParametrizedTConstantVDomain(phase=liq, initialconds=IC_TV, constantspecies=["oxygen"])
t0 = time.time(); f = gdom.dydt(gdom.y0, p, t=np.array([10.0])); print("dydt", time.time() - t0, np.abs(f).max(), flush=True)
t0 = time.time(); J = gdom.jacobiany(gdom.y0, p, t=np.array([10.0])); print("jac", time.time() - t0, np.abs(J).max(), flush=True)
for tend, ms in ((1.0, 2000), (60.0, 5000), (600.0, 20000), (4800.0, 50000)):
    t0 = time.time()
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p).ode, rb.RMSB200(max_steps=ms), abstol=1e-18, reltol=1e-9)
    print(tend, sol.retcode, sol.stats[0], f"{time.time() - t0:.2f} s", flush=True)
