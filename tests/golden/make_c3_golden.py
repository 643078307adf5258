"""Golden trajectory for config C3 at its stated size (tests/test_gpu_large.py): the octane-family mechanism of
tests/synth.py::write_octane_families (36 families, seed 20261019: 1015 species / 11 593 reactions), ConstantTV, T = 450 K, V = 1,
oxygen held constant (the reference's C3 charge, src/TestReactors.jl:18-22), integrated by SciPy BDF over the NumPy oracle
(oracle/rms_oracle.py: dydtreactor + analytic jacobiany as a sparse matrix) at rtol 1e-9 / atol 1e-18.
Writes tests/golden/c3_octane_families.npz (times, states, and the oracle's dydt / kf / krev at three perturbed states).
Run from the repo root:  python tests/golden/make_c3_golden.py   (11 CPU-minutes; rtol 1e-10 / 1e-11, also with the C++ restatement
supplying the callbacks, were stopped after 4 and 1 CPU-hours: the golden therefore carries SciPy's global error at rtol 1e-9, about
1.2e-6 on the species above 1e-9 of the largest -- see the bars in tests/test_gpu_large.py)"""
import sys
import time
from pathlib import Path

import numpy as np
from scipy.integrate import solve_ivp
from scipy.sparse import csc_matrix

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tests"):
    sys.path.insert(0, str(p))
import synth  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402

K = 36
path = "/tmp/rmsb_octane_families_golden.rms"
info = synth.write_octane_families(path, families=K)
d = ro.readinput(path)
ph = ro.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0])
ic = {"T": 450.0, "V": 1.0, **synth.octane_family_moles(K)}
dom = ro.ConstantTVDomain(ph, ic, constantspecies=("oxygen",))
tq = np.array([1.0e3, 1.0e4, 5.0e4, 1.4e5])
t0 = time.time()
sol = solve_ivp(lambda t, y: ro.dydtreactor(dom, y, t, dom.p), (0.0, tq[-1]), dom.y0, method="BDF", rtol=1e-9, atol=1e-18,
                jac=lambda t, y: csc_matrix(ro.jacobiany(dom, y, t, dom.p)), t_eval=tq)
assert sol.status == 0
print("SciPy BDF", time.time() - t0, "s", sol.nfev, sol.njev, sol.nlu)
rng = np.random.default_rng(1)
Ns = ph.Ns
ys = np.stack([dom.y0, np.abs(dom.y0 + 10.0 ** rng.uniform(-3, 1, Ns)), np.abs(dom.y0 * rng.uniform(0.5, 1.5, Ns) + 1e-6)])
f = np.stack([ro.dydtreactor(dom, y, 0.0, dom.p) for y in ys])
np.savez_compressed(ROOT / "tests" / "golden" / "c3_octane_families.npz", tq=tq, yq=sol.y.T, ys=ys, f=f, kfs=dom.kfs, krevs=dom.krevs,
                    ns=Ns, nr=ph.Nr, names=np.array(ph.names))
print("x(octane, family 0) at t_end:", sol.y[ph.names.index("F00_octane"), -1] / sol.y[:, -1].sum())
