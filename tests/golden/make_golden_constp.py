"""Generates tests/golden/constp_vectors.npz: oracle dydt / jacobiany / jacobianp of the ConstantPDomain (adiabatic, constant
pressure; src/Domain.jl:173-239,1141-1192,1950-1953,2456-2632,3433-3460) on the ethane mechanism at the reference's own
conditions (src/TestReactors.jl:344-370), plus a tight-tolerance trajectory.
Run in the build container:  python tests/golden/make_golden_constp.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402

IC = {"T": 1000.0, "P": 2.0e5, "ethane": 1.0, "Ar": 1.0, "O2": 3.5}

if __name__ == "__main__":
    d = ro.readinput(cases.fixture("ethane"))
    ig = ro.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"])
    dom = ro.ConstantPDomain(ig, IC)
    Ns = ig.Ns
    rng = np.random.default_rng([cases.SEED, 55])
    ys = []
    for _ in range(6):
        y = dom.y0.copy()
        tot = y[:Ns].sum()
        y[:Ns] = np.abs(y[:Ns] + tot * 10.0 ** rng.uniform(-8, -2, Ns) * rng.uniform(0.1, 1, Ns))
        y[Ns] *= rng.uniform(0.9, 1.6)
        y[Ns + 1] *= rng.uniform(0.8, 1.5)
        ys.append(y)
    ys = np.array(ys)
    pp = np.concatenate([rng.normal(0, 500.0, Ns), rng.uniform(0.5, 2.0, ig.Nr)])
    out = {"y": ys, "p": pp,
           "f": np.array([ro.dydtreactor(dom, y, 0.0, dom.p) for y in ys]),
           "J": np.array([ro.jacobiany(dom, y, 0.0, dom.p) for y in ys]),
           "Jp": np.array([ro.jacobianp(dom, y, 0.0, dom.p) for y in ys]),
           "f_p": np.array([ro.dydtreactor(dom, y, 0.0, pp) for y in ys]),
           "Jp_p": np.array([ro.jacobianp(dom, y, 0.0, pp) for y in ys])}
    tq = np.array([0.01, 0.05, 0.07, 0.0732, 0.0735, 0.08, 0.2, 1.0])
    r = co.CpuMech(ig).solve("ConstantP", dom.y0[None], dom.cond[None], 1.0, 1e-11, 1e-20, p=dom.p, tsave=tq, nthreads=1)
    assert r["status"][0] == 0
    out["tq"], out["yq"] = tq, r["ysave"][0]
    np.savez_compressed(Path(__file__).resolve().parent / "constp_vectors.npz", **out)
    print("written", {k: v.shape for k, v in out.items()})
