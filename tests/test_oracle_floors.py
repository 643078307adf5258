"""The cancellation floors of oracle/floors.py (used by tests/test_gpu_elementwise.py), exercised on the CPU: the C++ CPU port and the
NumPy oracle are two independent FP64 evaluations of the same formulas (different summation orders), so their plain elementwise
difference must be at rounding level on every entry above the floor and within a few floors below it."""
import numpy as np
import pytest

import cases
from oracle import cpu_oracle as co
from oracle import floors as fl
from oracle import rms_oracle as ro


@pytest.mark.parametrize("case", ["superminimal_V", "superminimal_TP", "ethane_V", "ethane_TP", "liquid_TV", "surface_TAPhi"])
def test_cpu_port_vs_numpy_oracle_elementwise(case):
    od = cases.oracle_domain(case)
    ys = cases.states(case, od.y0, od.phase.Ns, od.kind)[:2]
    cm = co.CpuMech(od.phase)
    po = od.p if od.kind in ("ConstantV", "ConstantTP") else None
    for y in ys:
        f, M = fl.dydt_terms(od, y, 0.0, po)
        assert np.all(M >= np.abs(f) * (1 - 1e-12))                  # a sum of magnitudes bounds the signed sum
        fc = cm.dydt(od.kind, y[None], od.cond[None], po, constspc=od.constantspeciesinds)[0]
        r = fl.report(fc, f, fl.EPS * M, 1e-12, slack=4.0)
        assert r["n_above"] >= 0.6 * r["n"] and r["max_rel_above"] < 1e-12 and r["worst_excess"] < 1.0, r
        Jo, Jf = fl.jac_floor(lambda yy, pp: ro.jacobiany(od, yy, 0.0, pp), od, y, po, nsample=4)
        Jc = cm.jac(od.kind, y[None], od.cond[None], po, constspc=od.constantspeciesinds)[0]
        r = fl.report(Jc, Jo, Jf, 1e-12)
        assert r["max_rel_above"] < 1e-12 and r["worst_excess"] < 1.0, r


def test_floor_flags_a_cancelling_entry():
    """An entry built as a difference of two nearly equal terms is below the floor; a plain product is above it."""
    def fn(u, _):
        return np.array([u[0] * u[1] - u[2] * u[3], u[0] * u[1]])
    u = np.array([1.0, 3.0, 1.5, 2.0 - 1e-13])
    ref, floor = fl.numeric_floor(fn, [u, None], nsample=8)
    rep = fl.report(ref * (1 + 3e-13), ref, floor, 1e-12)
    assert floor[0] > 1e-16 and floor[0] > 100 * abs(ref[0]) * 1e-13 and floor[1] <= 1e-13 * abs(ref[1]) * 100
    assert rep["n_above"] == 1 and rep["max_rel_above"] < 1e-12
