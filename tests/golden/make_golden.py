"""Generates tests/golden/parity_vectors.npz: oracle dydt / Jacobian at seeded states for every case in
tests/cases.py, plus the oracle's tight-tolerance trajectory values for the reference's golden points.
Run in the build container:  python tests/golden/make_golden.py   (needs numpy, scipy, yaml)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
from oracle import integrate as oi  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402

if __name__ == "__main__":
    out = {}
    for name in cases.CASES:
        dom = cases.oracle_domain(name)
        Ns = dom.phase.Ns
        ys = cases.states(name, dom.y0, Ns, dom.kind)
        p = dom.p if dom.kind in ("ConstantV", "ConstantTP") else None
        out[f"{name}/y"] = ys
        out[f"{name}/f"] = np.array([ro.dydtreactor(dom, y, 0.0, p) for y in ys])
        out[f"{name}/J"] = np.array([ro.jacobiany(dom, y, 0.0, p) for y in ys])
        if dom.kind == "ConstantV":
            pp = cases.parameter_vector(name, Ns, dom.phase.Nr)
            out[f"{name}/p"] = pp
            out[f"{name}/f_p"] = np.array([ro.dydtreactor(dom, y, 0.0, pp) for y in ys])
        print(name, ys.shape)
    # trajectory goldens from the oracle at tight tolerance (the reference's own pinned points)
    dom = cases.oracle_domain("superminimal_V")
    sol = oi.solve(dom, (0, 0.101), p=dom.p, rtol=1e-10, atol=1e-20)
    tq = np.array([1e-3, 1e-2, 0.03, 0.038, 0.0385, 0.039, 0.05, 0.1])
    out["superminimal_V/tq"] = tq
    out["superminimal_V/yq"] = np.array([sol.sol(t) for t in tq])
    out["superminimal_V/idt_grid"] = np.array(oi.idt_grid(sol.sol))
    dom = cases.oracle_domain("superminimal_TP")
    sol = oi.solve(dom, (0, 150.11094), p=dom.p, rtol=1e-10, atol=1e-20)
    tq = np.array([1.0, 20.44002454, 100.0, 150.11094])
    out["superminimal_TP/tq"] = tq
    out["superminimal_TP/yq"] = np.array([sol.sol(t) for t in tq])
    np.savez_compressed(Path(__file__).resolve().parent / "parity_vectors.npz", **out)
    print("written")
