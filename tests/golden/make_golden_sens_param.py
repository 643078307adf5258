"""Generates tests/golden/sens_param_vectors.npz: the oracle's parameter Jacobians and forward sensitivities of the Parametrized*
domains (time profiles) on the reference's test inputs: ParametrizedTP (src/TestReactors.jl:409-412) over the reference's own
tspan (0, 0.101), ParametrizedV / ParametrizedP (:427-454) over their induction periods to 0.02 s (the augmented 390-dimensional
system through the thermal runaway takes SciPy more than a quarter of an hour), and jacobianp at seeded states and times for all four Parametrized domains.
The augmented system [y; S] (oracle/sensitivity.py) is integrated by SciPy BDF; one column is spot-checked against central finite
differences of the state solution.  Run from the repo root:  python tests/golden/make_golden_sens_param.py"""
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

IC = {"TP": {"T": [1000.0, 1400.0, 1500.0], "ts": [0.0, 100.0, 200.0], "P": [1.0e5, 1.8e5, 2.0e5], "H2": 0.67, "O2": 0.33},
      "V": {"T": 1000.0, "ts": [0.0, 100.0, 200.0], "V": [0.01, 0.05, 0.1], "H2": 0.67, "O2": 0.33},
      "P": {"T": 1000.0, "ts": [0.0, 100.0, 200.0], "P": [2.0e5, 3.0e5, 5.0e5], "H2": 0.67, "O2": 0.33},
      "TV": {"ts": [0.0, 600.0, 1200.0], "T": [450.0, 490.0, 500.0], "V": 1.0, "octane": 6154.0, "oxygen": 4.953}}


def domains():
    d = ro.readinput(cases.fixture("superminimal"))
    ig = ro.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"])
    od = ro.readinput(cases.fixture("liquid_phase"))
    liq = ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0])
    return {"TP": ro.ParametrizedTPDomain(ig, IC["TP"]), "V": ro.ParametrizedVDomain(ig, IC["V"]), "P": ro.ParametrizedPDomain(ig, IC["P"]),
            "TV": ro.ParametrizedTConstantVDomain(liq, IC["TV"], constantspecies=("oxygen",))}


STATES = {"TP": ("superminimal_TP", "ConstantTP", [0.1, 50.0, 180.0]), "V": ("superminimal_V", "ConstantV", [0.1, 37.0, 150.0]),
          "P": ("superminimal_V", "ConstantV", [0.1, 37.0, 150.0]), "TV": ("liquid_TV", "ConstantTV", [0.0, 700.0, 32977.61568])}

if __name__ == "__main__":
    out = {}
    doms = domains()
    for k, dom in doms.items():
        case, kind, ts = STATES[k]
        ys = cases.states(case, dom.y0, dom.phase.Ns, kind)[:3]
        out[f"{k}/y"], out[f"{k}/t"] = ys, np.array(ts)
        out[f"{k}/Jp"] = np.array([ro.jacobianp(dom, y, t, dom.p) for y, t in zip(ys, ts)])
    np.savez_compressed(Path(__file__).resolve().parent / "sens_param_vectors.npz", **out)
    for k, tend, tq, rtol in (("P", 0.02, np.array([0.01, 0.02]), 1e-10), ("V", 0.02, np.array([0.01, 0.02]), 1e-10), ("TP", 0.101, np.array([0.05, 0.101]), 1e-10)):
        dom = doms[k]
        t0 = time.time()
        t, y, S = osens.solve_forward(dom, (0.0, tend), dom.p, rtol=rtol, atol=1e-22, t_eval=tq, exact_jac=False)
        print(k, "augmented solve", round(time.time() - t0), "s", flush=True)
        out[f"{k}/tq"], out[f"{k}/yq"], out[f"{k}/Sq"] = tq, y, S
        Ns = dom.phase.Ns
        norms = np.abs(S[-1]).max(axis=1)
        j = Ns + int(np.argmax(norms[Ns:]))
        fd = osens.finite_difference(dom, (0.0, tend), dom.p, j, rtol=1e-9, atol=1e-20, t_eval=tq)
        err = np.abs(fd - S[:, j, :]).max() / np.abs(S[:, j, :]).max()
        print(k, "FD check param", j, "rel err", err, flush=True)
        assert err < 5e-3
        np.savez_compressed(Path(__file__).resolve().parent / "sens_param_vectors.npz", **out)
    print("written")
