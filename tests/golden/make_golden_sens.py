"""Generates tests/golden/sens_vectors.npz: oracle jacobianp at the seeded states of tests/cases.py and the oracle's
forward sensitivities for config C5 (superminimal ConstantTP, src/TestReactors.jl:191-255) plus a ConstantV case.
Run in the build container:  python tests/golden/make_golden_sens.py
"""
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

if __name__ == "__main__":
    out = {}
    for name in cases.CASES:
        dom = cases.oracle_domain(name)
        Ns = dom.phase.Ns
        ys = cases.states(name, dom.y0, Ns, dom.kind)
        out[f"{name}/y"] = ys
        out[f"{name}/Jp"] = np.array([ro.jacobianp(dom, y, 0.0, dom.p) for y in ys])
        if dom.kind == "ConstantV":
            pp = cases.parameter_vector(name, Ns, dom.phase.Nr)
            out[f"{name}/p"] = pp
            out[f"{name}/Jp_p"] = np.array([ro.jacobianp(dom, y, 0.0, pp) for y in ys])
        print(name, out[f"{name}/Jp"].shape, flush=True)
    # C5: forward sensitivities of the superminimal ConstantTP reactor
    dom = cases.oracle_domain("superminimal_TP")
    tq = np.array([1.0, 20.44002454, 100.0, 150.11094])
    t0 = time.time()
    t, y, S = osens.solve_forward(dom, (0.0, 150.11094), dom.p, rtol=1e-9, atol=1e-22, t_eval=tq, exact_jac=False)
    print("C5 augmented solve", time.time() - t0, "s", flush=True)
    out["superminimal_TP/tq"], out["superminimal_TP/yq"], out["superminimal_TP/Sq"] = tq, y, S
    Ns = dom.phase.Ns
    norms = np.abs(S[-1]).max(axis=1)
    picks = (int(np.argmax(norms[:Ns])), Ns + int(np.argmax(norms[Ns:])), Ns + int(np.argsort(norms[Ns:])[-3]))
    for j in picks[1:2]:                        # spot-check against finite differences (independent of jacobianp)
        fd = osens.finite_difference(dom, (0.0, 150.11094), dom.p, j, rtol=1e-9, atol=1e-20, t_eval=tq)
        err = np.abs(fd - S[:, j, :]).max() / np.abs(S[:, j, :]).max()
        print("FD check param", j, "rel err", err, flush=True)
        assert err < 2e-3
    np.savez_compressed(Path(__file__).resolve().parent / "sens_vectors.npz", **out)
    # ConstantV (variable k, T and P rows): the induction period before ignition (the augmented 420-dimensional system
    # through the ignition front takes SciPy hours; the GPU test checks the post-ignition values against finite differences)
    dom = cases.oracle_domain("superminimal_V")
    tq = np.array([0.01, 0.02, 0.03])
    t0 = time.time()
    t, y, S = osens.solve_forward(dom, (0.0, 0.03), dom.p, rtol=1e-8, atol=1e-22, t_eval=tq, exact_jac=False)
    print("V augmented solve", time.time() - t0, "s", flush=True)
    out["superminimal_V/tq"], out["superminimal_V/yq"], out["superminimal_V/Sq"] = tq, y, S
    np.savez_compressed(Path(__file__).resolve().parent / "sens_vectors.npz", **out)
    print("written")
