"""Boundary interfaces: Inlet, Outlet, VolumetricFlowRateInlet / Outlet, VolumeMaintainingOutlet in dydt, jacobiany and
solve on the GPU against the oracle's restatement of src/Domain.jl:1839-1998 (calcdomainderivatives!) and :2220-2294, 2322-2448,
2493-2628, 3057-3088 (jacobiany!), through the reference's own call sequence (src/TestReactors.jl:259-296, 41-63)."""
import numpy as np
import pytest

import cases
import rmsb200 as rb
from oracle import floors as fl
from oracle import integrate as oi
from oracle import rms_oracle as ro
from rmsb200 import _lib as L

pytestmark = pytest.mark.gpu

SETS = {
    # case -> list of (class name, args) built for both sides
    "superminimal_TP": [("Inlet", ({"H2": 0.67, "O2": 0.33, "T": 1000.0, "P": 1e5}, 0.001)), ("Outlet", (0.001,))],
    "superminimal_V": [("Inlet", ({"H2": 0.67, "O2": 0.33, "T": 900.0, "P": 1e5}, 0.002)), ("Outlet", (0.001,)),
                       ("VolumetricFlowRateInlet", ({"H2": 8.0, "O2": 4.0, "T": 800.0, "P": 1e5}, 1e-4)), ("VolumetricFlowRateOutlet", (2e-4,))],
    "ethane_V": [("Inlet", ({"ethane": 0.3, "O2": 0.7, "T": 900.0, "P": 1e5}, 0.002)), ("Outlet", (0.001,))],
    "ethane_TP": [("Inlet", ({"ethane": 0.3, "O2": 0.7, "T": 900.0, "P": 1e5}, 0.002)), ("VolumetricFlowRateOutlet", (1e-3,)),
                  ("VolumeMaintainingOutlet", ())],
    "liquid_TV": [("VolumetricFlowRateInlet", ({"octane": 6154.0, "oxygen": 4.953, "T": 450.0, "P": 1e5}, 1.0)), ("VolumetricFlowRateOutlet", (1.0,)),
                  ("Outlet", (3.0,))],
    "surface_TAPhi": [("Outlet", (1e-9,))],
}


def _build(case):
    odom = cases.oracle_domain(case)
    gdom, y0, p = cases.product_domain(case)
    oifs = [getattr(ro, nm)(odom, *args) for nm, args in SETS[case]]
    gifs = [getattr(rb, nm)(gdom, *args) for nm, args in SETS[case]]
    return odom, gdom, y0, p, oifs, gifs


@pytest.mark.parametrize("case", list(SETS))
def test_interfaces_dydt_and_jacobian_parity(case):
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    odom, gdom, y0, p, oifs, gifs = _build(case)
    Ns = odom.phase.Ns
    ys = cases.states(case, odom.y0, Ns, odom.kind)[:3]
    po = odom.p if odom.kind in ("ConstantV", "ConstantTP") else None
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.constantspeciesinds = gdom.constantspeciesinds
    ens.p, ens._p_default = gdom.p, gdom._p_default
    row = rb.pack_interfaces(gdom, gifs)
    f, J = ens.dydt(ys, p, interfaces=row), ens.jacobiany(ys, p, interfaces=row)
    f0 = ens.dydt(ys, p)
    assert not np.array_equal(f, f0)                                          # the interfaces do something
    for k, y in enumerate(ys):
        fo = ro.dydtreactor(odom, y, 0.0, po, oifs)
        _, M = fl.dydt_terms(odom, y, 0.0, po)                                # chemistry floor; the flow terms add no cancellation of their own
        rep = fl.report(f[k], fo, fl.EPS * (M + np.abs(fo)), 1e-12, slack=4.0)
        assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, (case, k, rep)
        Jo, Jf = fl.jac_floor(lambda yy, pp: ro.jacobiany(odom, yy, 0.0, pp, interfaces=oifs), odom, y, po)
        rep = fl.report(J[k], Jo, Jf, 1e-12)
        assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, (case, k, [rep[q] for q in ("max_rel_above", "worst_excess", "worst_at", "worst_x", "worst_ref", "worst_floor")])


def test_interfaces_constant_p_with_volume_maintaining_outlet():
    d = rb.readinput(cases.fixture("ethane")); ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
    od = ro.readinput(cases.fixture("ethane")); oig = ro.IdealGas(od["phase"]["Species"], od["phase"]["Reactions"])
    ic = {"T": 1200.0, "P": 2e5, "ethane": 0.3, "O2": 0.6, "Ar": 0.1}
    gdom, y0, p = rb.ConstantPDomain(phase=ig, initialconds=ic)
    odom = ro.ConstantPDomain(oig, ic)
    spec = [("Inlet", ({"ethane": 0.3, "O2": 0.7, "T": 900.0, "P": 1e5}, 0.002)), ("Outlet", (0.001,)),
            ("VolumetricFlowRateInlet", ({"ethane": 3.0, "O2": 7.0, "T": 800.0, "P": 1e5}, 1e-4)), ("VolumetricFlowRateOutlet", (2e-4,))]
    for with_vm in (False, True):
        sp = spec + ([("VolumeMaintainingOutlet", ())] if with_vm else [])
        oifs = [getattr(ro, nm)(odom, *a) for nm, a in sp]
        gifs = [getattr(rb, nm)(gdom, *a) for nm, a in sp]
        ys = cases.states("ethane_V", odom.y0, ig.Ns, "ConstantV")[:2]
        ens = rb.Domain(gdom.kind, ig, ys, np.repeat(gdom.cond, len(ys), axis=0))
        row = rb.pack_interfaces(gdom, gifs)                                       # packed against the domain the interfaces belong to
        f, J = ens.dydt(ys, p, interfaces=row), ens.jacobiany(ys, p, interfaces=row)
        for k, y in enumerate(ys):
            fo = ro.dydtreactor(odom, y, 0.0, odom.p, oifs)
            _, M = fl.dydt_terms(odom, y, 0.0, odom.p)
            rep = fl.report(f[k], fo, fl.EPS * (M + np.abs(fo)), 1e-12, slack=4.0)
            assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, (with_vm, k, rep)
            if with_vm:
                assert f[k][ig.Ns + 1] == 0.0                                      # the outlet holds the volume
            Jo, Jf = fl.jac_floor(lambda yy, pp: ro.jacobiany(odom, yy, 0.0, pp, interfaces=oifs), odom, y, odom.p)
            rep = fl.report(J[k], Jo, Jf, 1e-12)
            assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, (with_vm, k, rep["worst_at"], rep["worst_excess"])


def test_cstr_trajectory_reference_call_sequence():
    """src/TestReactors.jl:259-296: ConstantTP H2/O2 reactor with an Inlet and an Outlet of 1e-3 mol/s to t = 150.11094 s, state at
    t = 20.44002454 s against SciPy BDF over the oracle (the reference pins only Jacobian consistency and sensitivities there)."""
    odom, gdom, y0, p, oifs, gifs = _build("superminimal_TP")
    tq = np.array([20.44002454, 150.11094])
    react = rb.Reactor(gdom, y0, (0.0, 150.11094), gifs, p=p)
    sol = rb.solve(react.ode, abstol=1e-20, reltol=1e-10, saveat=tq)
    assert sol.retcode == ["Success"]
    from scipy.integrate import solve_ivp
    ref = solve_ivp(lambda t, y: ro.dydtreactor(odom, y, t, odom.p, oifs), (0.0, tq[-1]), odom.y0, method="BDF", rtol=1e-11, atol=1e-22,
                    jac=lambda t, y: ro.jacobiany(odom, y, t, odom.p, interfaces=oifs), t_eval=tq)
    assert ref.status == 0
    for k in range(2):
        yr = ref.y[:, k]
        big = np.abs(yr) > 1e-10 * np.abs(yr).max()
        assert np.max(np.abs(sol.u[0, k] - yr)[big] / np.abs(yr)[big]) < 1e-6
    closed = rb.solve(rb.Reactor(gdom, y0, (0.0, 150.11094), p=p).ode, abstol=1e-20, reltol=1e-10)
    iH2 = gdom.phase.names.index("H2")
    assert abs(closed.yend[0, iH2] - sol.yend[0, iH2]) > 1e-6 * abs(sol.yend[0, iH2])   # the flow changes the answer


def test_residence_time_sweep_per_reactor_rows():
    """A CSTR ensemble: every reactor its own inlet / outlet flow (rows of the interface matrix).  Each reactor agrees with the same
    reactor solved alone with a shared row; long residence times approach the closed reactor."""
    odom, gdom, y0, p, oifs, gifs = _build("superminimal_V")
    n = 64
    rng = np.random.default_rng(3)
    flows = np.zeros((n, 4))
    flows[:, 0] = 10.0 ** rng.uniform(-6, -1, n)
    flows[:, 1] = flows[:, 0]
    flows[0, :2] = 0.0
    ens = rb.Domain(gdom.kind, gdom.phase, np.repeat(gdom.y0, n, axis=0), np.repeat(gdom.cond, n, axis=0))
    ens.p, ens._p_default = gdom.p, gdom._p_default
    rows = rb.pack_interfaces(gdom, gifs[:2], flows)
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.05), rows, p=p).ode, abstol=1e-16, reltol=1e-8)
    assert np.all(sol.status == 0)
    for k in (1, 17, n - 1):
        one = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.05), rows[k], p=p).ode, abstol=1e-16, reltol=1e-8)
        assert np.array_equal(one.yend[0], sol.yend[k])
    closed = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.05), p=p).ode, abstol=1e-16, reltol=1e-8)
    np.testing.assert_allclose(sol.yend[0], closed.yend[0], rtol=1e-6, atol=1e-30)
    slowest = int(np.argmin(flows[1:, 0])) + 1
    assert abs(sol.yend[slowest, 12] - closed.yend[0, 12]) < 1e-2 * closed.yend[0, 12]
    with pytest.raises(NotImplementedError):
        rb.Inlet(gdom, {"H2": 1.0, "T": 900.0, "P": 1e5}, lambda t: 1e-3 * (1 + t))    # time-dependent flows are not built
