"""Generates tests/golden/sens_if_vectors.npz: the oracle's forward sensitivities and parameter Jacobians WITH boundary interfaces.
  * the reference's own case (src/TestReactors.jl:259-296): superminimal ConstantTP, Inlet + Outlet at F = 0.001 mol/s, t_end = 150.11094;
  * a ConstantV case with an Inlet (the only configuration whose jacobianp! gains interface terms, src/Domain.jl:3421-3430), to 0.02 s.
The augmented system [y; S] (oracle/sensitivity.py) is integrated by SciPy BDF; one column is spot-checked against central finite
differences of the state solution, which do not use jacobianp.  Run from the repo root:  python tests/golden/make_golden_sens_if.py"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402
from oracle import sensitivity as osens  # noqa: E402

IF_TP = [("Inlet", ({"H2": 0.67, "O2": 0.33, "T": 1000.0, "P": 1e5}, 0.001)), ("Outlet", (0.001,))]
IF_V = [("Inlet", ({"H2": 0.67, "O2": 0.33, "T": 900.0, "P": 1e5}, 0.002)), ("Outlet", (0.001,))]

if __name__ == "__main__":
    out = {}
    for case, spec, tend, tq, rtol in (("superminimal_TP", IF_TP, 150.11094, np.array([20.44002454, 100.0, 150.11094]), 1e-9),
                                       ("superminimal_V", IF_V, 0.02, np.array([0.01, 0.02]), 1e-8)):
        dom = cases.oracle_domain(case)
        ifs = [getattr(ro, nm)(dom, *args) for nm, args in spec]
        Ns = dom.phase.Ns
        ys = cases.states(case, dom.y0, Ns, dom.kind)[:3]
        out[f"{case}/y"] = ys
        out[f"{case}/Jp"] = np.array([ro.jacobianp(dom, y, 0.0, dom.p, ifs) for y in ys])
        t0 = time.time()
        t, y, S = osens.solve_forward(dom, (0.0, tend), dom.p, rtol=rtol, atol=1e-22, t_eval=tq, exact_jac=False, interfaces=ifs)
        print(case, "augmented solve", round(time.time() - t0), "s", flush=True)
        out[f"{case}/tq"], out[f"{case}/yq"], out[f"{case}/Sq"] = tq, y, S
        norms = np.abs(S[-1]).max(axis=1)
        j = Ns + int(np.argmax(norms[Ns:]))
        fd = osens.finite_difference(dom, (0.0, tend), dom.p, j, rtol=1e-9, atol=1e-20, t_eval=tq, interfaces=ifs)
        err = np.abs(fd - S[:, j, :]).max() / np.abs(S[:, j, :]).max()
        print(case, "FD check param", j, "rel err", err, flush=True)
        assert err < 5e-3
        np.savez_compressed(Path(__file__).resolve().parent / "sens_if_vectors.npz", **out)
    print("written")
