"""Golden for the SciPy check of the headline sweep (tests/test_gpu_parity.py::test_c2_subsample_vs_scipy): the first 64 conditions of
the C2 (T0, P0, phi) sweep of SURVEY 8(d) -- the same generator as tests/test_gpu_parity.py::sweep -- integrated to t = 0.1 s by SciPy
BDF over the NumPy oracle (oracle/integrate.py: dydtreactor + analytic jacobiany), rtol 1e-10 / atol 1e-20: final state and ignition
delay (time of max dT/dt over the accepted steps, parabola-refined: oracle/integrate.py::idt_continuous).
Independent of the C++ port and of the GPU integrator.  Writes tests/golden/c2_scipy_subsample.npz.
Run from the repo root:  python tests/golden/make_c2_scipy_subsample.py   (8 processes, a few minutes)"""
import sys
import time
from multiprocessing import Pool
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
from oracle import integrate as oi  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402

N = 64
R = 8.314


def conditions():
    rng = np.random.default_rng(cases.SEED)          # tests/test_gpu_parity.py::sweep
    d = ro.readinput(cases.fixture("superminimal"))
    ig = ro.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"])
    T0, P0, phi = rng.uniform(900, 1500, N), 10 ** rng.uniform(5, 6.7, N), rng.uniform(0.5, 2.0, N)
    ns = np.zeros((N, ig.Ns))
    ns[:, ig.spcdict["H2"]] = 2 * phi
    ns[:, ig.spcdict["O2"]] = 1.0
    ns[:, ig.spcdict["N2"]] = 3.76 * rng.integers(0, 2, N)
    ns /= ns.sum(axis=1, keepdims=True)
    return ig, T0, P0, ns


def one(k):
    ig, T0, P0, ns = conditions()
    ic = {"T": float(T0[k]), "P": float(P0[k])}
    ic.update({nm: float(ns[k, i]) for i, nm in enumerate(ig.names) if ns[k, i] > 0.0})
    dom = ro.ConstantVDomain(ig, ic)
    sol = oi.solve(dom, (0.0, 0.1), p=dom.p, rtol=1e-10, atol=1e-20)
    assert sol.status == 0
    return k, dom.y0, dom.cond, sol.y[:, -1], oi.idt_continuous(sol, ig.Ns, dom, dom.p), sol.nfev


if __name__ == "__main__":
    t0 = time.time()
    with Pool(8) as pool:
        res = sorted(pool.map(one, range(N)))
    y0 = np.array([r[1] for r in res]); cond = np.array([r[2] for r in res]); yend = np.array([r[3] for r in res])
    idt = np.array([r[4] for r in res])
    print(f"{N} conditions, {sum(r[5] for r in res)} oracle evaluations, {time.time() - t0:.0f} s")
    np.savez_compressed(ROOT / "tests" / "golden" / "c2_scipy_subsample.npz", y0=y0, cond=cond, yend=yend, idt=idt, rtol=1e-10, atol=1e-20)
    print("written")
