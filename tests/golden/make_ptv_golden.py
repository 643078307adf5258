"""Golden trajectory for ParametrizedTConstantVDomain (tests/test_gpu_parametrized.py): the reference's test charge
(src/TestReactors.jl:148-153: liquid octane mechanism, T ramp 450 -> 490 -> 500 K over 1200 s, V = 1, oxygen held constant)
integrated by SciPy BDF over the NumPy oracle (oracle/rms_oracle.py: dydtreactor + analytic jacobiany) to t = 4800 s.
Writes tests/golden/ptv_liquid.npz.  Run from the repo root:  python tests/golden/make_ptv_golden.py [rtol]
(rtol 1e-9 / atol 1e-20: 37 000 oracle evaluations, 3 CPU-minutes)"""
import sys
import time
from pathlib import Path

import numpy as np
from scipy.integrate import solve_ivp

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402

rtol = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-9
od = ro.readinput(cases.fixture("liquid_phase"))
liq = ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0])
ic = {"ts": [0.0, 600.0, 1200.0], "T": [450.0, 490.0, 500.0], "V": 1.0, "octane": 6154.0, "oxygen": 4.953}
dom = ro.ParametrizedTConstantVDomain(liq, ic, constantspecies=("oxygen",))
tq = np.array([600.0, 1200.0, 4800.0])
t0 = time.time()
ref = solve_ivp(lambda t, y: ro.dydtreactor(dom, y, t, dom.p), (0.0, tq[-1]), dom.y0, method="BDF", rtol=rtol, atol=1e-20,
                jac=lambda t, y: ro.jacobiany(dom, y, t, dom.p), t_eval=tq)
assert ref.status == 0
print(f"rtol {rtol:g}: {ref.nfev} rhs, {ref.njev} jac, {time.time() - t0:.0f} s")
out = ROOT / "tests" / "golden" / ("ptv_liquid.npz" if len(sys.argv) < 3 else sys.argv[2])
np.savez(out, tq=tq, yq=ref.y.T.copy(), rtol=rtol, atol=1e-20)
print("wrote", out)
