"""Shared parity-case definitions: the same seeded inputs feed the oracle (golden generator, CPU tests)
and the CUDA path (GPU tests).  Cases follow SURVEY.md section 8(d) / BASELINE.json configs."""
from __future__ import annotations

from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
FX = ROOT / "tests" / "fixtures"
SEED = 20261019

A_CAT = 1e5 * 8.314 * 800.0 / 1e5        # src/TestReactors.jl:545-549

# name -> (fixture, phase key, phase kind, domain kind, initial conditions, constant species)
CASES = {
    "superminimal_V": ("superminimal", "phase", "gas", "ConstantV", {"T": 1000.0, "P": 10e5, "H2": 0.67, "O2": 0.33}, ()),
    "superminimal_TP": ("superminimal", "phase", "gas", "ConstantTP", {"T": 1000.0, "P": 1e5, "H2": 0.67, "O2": 0.33}, ()),
    "minimal_V": ("minimal", "phase", "gas", "ConstantV", {"T": 1100.0, "P": 2e5, "H2": 0.5, "O2": 0.4, "N2": 0.1}, ()),
    "ethane_V": ("ethane", "phase", "gas", "ConstantV", {"T": 1200.0, "P": 2e5, "ethane": 0.3, "O2": 0.6, "Ar": 0.1}, ()),
    "ethane_TP": ("ethane", "phase", "gas", "ConstantTP", {"T": 1200.0, "P": 2e5, "ethane": 0.3, "O2": 0.6, "Ar": 0.1}, ("Ar",)),
    "propane_V": ("propane", "phase", "gas", "ConstantV", {"T": 1100.0, "P": 3e5, "propane": 0.3, "O2": 0.6, "N2": 0.1}, ()),
    "liquid_TV": ("liquid_phase", "phase", "liquid", "ConstantTV", {"T": 450.0, "V": 1.0, "octane": 6154.0, "oxygen": 4.953}, ("oxygen",)),
    "surface_TAPhi": ("ch4o2cat", "surface", "surface", "ConstantTAPhi", {"T": 800.0, "A": A_CAT, "vacantX": 2.486e-5 * A_CAT}, ()),
}
NSTATE = {"propane_V": 2, "liquid_TV": 3, "surface_TAPhi": 3}


def fixture(name: str) -> str:
    return str(FX / f"{name}.rms")


def oracle_domain(case: str):
    from oracle import rms_oracle as ro
    fx, key, pk, dk, ic, cs = CASES[case]
    d = ro.readinput(fixture(fx))
    sp, rx = d[key]["Species"], d[key]["Reactions"]
    if pk == "gas":
        ph = ro.IdealGas(sp, rx)
    elif pk == "liquid":
        ph = ro.IdealDiluteSolution(sp, rx, d["Solvents"][0])
    else:
        ph = ro.IdealSurface(sp, rx, 2.486e-5)
    ic = {k: v for k, v in ic.items() if k in ("T", "P", "V", "A") or k in ph.spcdict}
    return getattr(ro, dk + "Domain")(ph, ic, constantspecies=cs)


def product_domain(case: str):
    import rmsb200 as rb
    fx, key, pk, dk, ic, cs = CASES[case]
    d = rb.readinput(fixture(fx))
    sp, rx = d[key]["Species"], d[key]["Reactions"]
    if pk == "gas":
        ph = rb.IdealGas(sp, rx, name=key)
    elif pk == "liquid":
        ph = rb.IdealDiluteSolution(sp, rx, d["Solvents"][0], name=key)
    else:
        ph = rb.IdealSurface(sp, rx, 2.486e-5, name=key)
    ic = {k: v for k, v in ic.items() if k in ("T", "P", "V", "A") or k in ph.spcdict}
    return getattr(rb, dk + "Domain")(phase=ph, initialconds=ic, constantspecies=list(cs))


def states(case: str, y0: np.ndarray, Ns: int, kind: str) -> np.ndarray:
    """Seeded perturbed states around y0: every species populated over 6 decades so all reactions fire."""
    rng = np.random.default_rng([SEED, sum(map(ord, case))])
    n = NSTATE.get(case, 6)
    ys = []
    for _ in range(n):
        y = y0.copy()
        tot = y[:Ns].sum()
        y[:Ns] = np.abs(y[:Ns] + tot * 10.0 ** rng.uniform(-8, -2, Ns) * rng.uniform(0.1, 1, Ns))
        if kind == "ConstantV":
            y[Ns] *= rng.uniform(0.9, 1.6)
            y[Ns + 1] *= rng.uniform(0.8, 1.5)
        ys.append(y)
    return np.array(ys)


def parameter_vector(case: str, Ns: int, Nr: int) -> np.ndarray:
    """Non-default ConstantV parameters p = [dG; f] (src/Domain.jl:1100,1112)."""
    rng = np.random.default_rng([SEED, 7, sum(map(ord, case))])
    return np.concatenate([rng.normal(0, 500.0, Ns), rng.uniform(0.5, 2.0, Nr)])


def err_dydt(f, f_ref):
    """max |f - f_ref| / (|f_ref| + 1e-6*max|f_ref|)"""
    return float(np.max(np.abs(f - f_ref) / (np.abs(f_ref) + 1e-6 * np.max(np.abs(f_ref)) + 1e-300)))


def err_jac(J, J_ref):
    """max |J - J_ref| / (|J_ref| + 1e-3 * column max|J_ref|): elementwise on every entry that is not
    pure cancellation noise relative to its column."""
    col = np.max(np.abs(J_ref), axis=0, keepdims=True)
    return float(np.max(np.abs(J - J_ref) / (np.abs(J_ref) + 1e-3 * col + 1e-300)))
