"""Generates tests/golden/coupled_vectors.npz: oracle dydt / Jacobian / trajectory values of the C4 coupled gas-surface
system (oracle/coupled.py).  Run in the build container:  python tests/golden/make_golden_coupled.py"""
import sys
from pathlib import Path

import numpy as np
from scipy.integrate import solve_ivp

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402
from oracle import coupled  # noqa: E402

if __name__ == "__main__":
    out = {}
    s = coupled.ch4o2cat_system(cases.fixture("ch4o2cat"))
    sol = solve_ivp(lambda t, y: s.dydt(y), (0, 0.1), s.y0, method="BDF", rtol=1e-10, atol=1e-22, jac=lambda t, y: s.jacobian(y), dense_output=True)
    assert sol.status == 0
    tq = np.array([5e-6, 1e-4, 1e-2, 0.1])
    out["tq"], out["yq"] = tq, np.array([sol.sol(t) for t in tq])
    ys = np.array([s.y0, sol.sol(5e-6), sol.sol(1e-3), sol.sol(0.1)])
    out["y"] = ys
    out["f"] = np.array([s.dydt(y) for y in ys])
    out["J"] = np.array([s.jacobian(y) for y in ys])
    out["fscale"] = np.array([s.term_scale(y) for y in ys])
    out["kf"] = np.concatenate([s.gas.kfs, s.surf.kfs, s.ikfs])
    out["kr"] = np.concatenate([s.gas.krevs, s.surf.krevs, s.ikrevs])
    # two more conditions of the C4 sweep: T and x(CH4) varied
    for k, (T, x) in enumerate([(700.0, 0.05), (1000.0, 0.2)]):
        sk = coupled.ch4o2cat_system(cases.fixture("ch4o2cat"), T=T, xCH4=x)
        so = solve_ivp(lambda t, y: sk.dydt(y), (0, 0.1), sk.y0, method="BDF", rtol=1e-10, atol=1e-22, jac=lambda t, y: sk.jacobian(y), dense_output=True)
        out[f"sweep{k}/cond"] = np.array([T, x]); out[f"sweep{k}/y0"] = sk.y0; out[f"sweep{k}/A"] = np.array(sk.A)
        out[f"sweep{k}/yq"] = np.array([so.sol(t) for t in tq])
    np.savez_compressed(Path(__file__).resolve().parent / "coupled_vectors.npz", **out)
    print("written", {k: v.shape for k, v in out.items()})
