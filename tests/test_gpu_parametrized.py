"""ParametrizedVDomain / ParametrizedPDomain: adiabatic reactors whose volume / pressure follows a time profile
(rapid-compression-machine and shock-tube traces), src/Domain.jl:331-498,1219-1368,2002-2095,2726-2748,2874-2896.  The reference's own
tests pin only analytic-vs-AD Jacobian consistency (src/TestReactors.jl:425-467); here the CUDA path is checked against the oracle's
restatement at 1e-12 (with the derived floors) on the reference's test inputs, and a compression trajectory against SciPy BDF."""
import numpy as np
import pytest

import cases
import rmsb200 as rb
from oracle import floors as fl
from oracle import rms_oracle as ro
from rmsb200 import _lib as L

pytestmark = pytest.mark.gpu

IC_V = {"T": 1000.0, "ts": [0.0, 100.0, 200.0], "V": [0.01, 0.05, 0.1], "H2": 0.67, "O2": 0.33}          # src/TestReactors.jl:427
IC_P = {"T": 1000.0, "ts": [0.0, 100.0, 200.0], "P": [2.0e5, 3.0e5, 5.0e5], "H2": 0.67, "O2": 0.33}      # :449


def _phases():
    d = rb.readinput(cases.fixture("superminimal")); ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
    od = ro.readinput(cases.fixture("superminimal")); oig = ro.IdealGas(od["phase"]["Species"], od["phase"]["Reactions"])
    return ig, oig


@pytest.mark.parametrize("which", ["V", "P"])
def test_parametrized_dydt_jacobian_parity(which):
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    ig, oig = _phases()
    if which == "V":
        gdom, y0, p = rb.ParametrizedVDomain(phase=ig, initialconds=IC_V)
        odom = ro.ParametrizedVDomain(oig, IC_V)
    else:
        gdom, y0, p = rb.ParametrizedPDomain(phase=ig, initialconds=IC_P)
        odom = ro.ParametrizedPDomain(oig, IC_P)
    np.testing.assert_allclose(y0, odom.y0, rtol=1e-14)
    ys = cases.states("superminimal_V", odom.y0, ig.Ns, "ConstantV")[:3]
    ts = np.array([0.1, 37.0, 150.0])                           # the reference's test time, and two well inside the profile
    par = cases.parameter_vector("superminimal_V", ig.Ns, ig.Nr)
    ens = rb.Domain(gdom.kind, ig, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.profile = gdom.profile
    ens.p, ens._p_default = gdom.p, gdom._p_default
    for pp in (p, par):
        f, J = ens.dydt(ys, pp, t=ts), ens.jacobiany(ys, pp, t=ts)
        for k, y in enumerate(ys):
            fo, M = fl.dydt_terms(odom, y, ts[k], pp)
            rep = fl.report(f[k], fo, fl.EPS * (M + np.abs(fo)), 1e-12, slack=4.0)
            assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, (which, k, rep)
            Jo, Jf = fl.jac_floor(lambda yy, qq: ro.jacobiany(odom, yy, ts[k], qq), odom, y, pp, t=ts[k])
            rep = fl.report(J[k], Jo, Jf, 1e-12)
            assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, (which, k, rep["worst_at"], rep["worst_excess"], rep["max_rel_above"])
            # the reference's criterion: analytic == derivative of the full RHS
            Jc = ro.jacobiany_complexstep(odom, y, ts[k], pp)
            assert not np.any(np.abs(J[k] - Jc) > 1e-4 * np.abs(Jc) + 1e-16 * np.abs(Jc).max())
    # the profile matters: at t = 150 the volume / pressure is far from its initial value
    plain = rb.Domain(gdom.kind, ig, ys, np.repeat(gdom.cond, len(ys), axis=0))
    plain.p, plain._p_default = gdom.p, gdom._p_default
    assert not np.allclose(plain.dydt(ys, p)[2], ens.dydt(ys, p, t=ts)[2], rtol=1e-3)


def test_rapid_compression_trajectory_vs_scipy():
    """A compression stroke: V(t) falls by 8x in 20 ms, H2/O2/N2 ignites near top dead centre.  Trajectory against SciPy BDF over
    the oracle, ignition delay against the oracle's dT/dt maximum; an ensemble of compression ratios through per-reactor profiles."""
    ig, oig = _phases()
    ts = np.array([0.0, 0.004, 0.008, 0.012, 0.016, 0.02, 0.03, 0.06])
    V0 = 1.0e-3
    shape = np.array([1.0, 0.93, 0.74, 0.48, 0.24, 0.125, 0.125, 0.125])
    ic = {"T": 420.0, "P": 1.0e5, "ts": ts, "V": V0 * shape, "H2": 2.0, "O2": 1.0, "N2": 3.76}
    n0 = 1.0e5 * V0 / (8.314 * 420.0)
    for k in ("H2", "O2", "N2"):
        ic[k] = ic[k] / 6.76 * n0
    ic.pop("P")
    gdom, y0, p = rb.ParametrizedVDomain(phase=ig, initialconds=ic)
    odom = ro.ParametrizedVDomain(oig, ic)
    tq = np.array([0.01, 0.02, 0.05])
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.05), p=p).ode, abstol=1e-18, reltol=1e-9, saveat=tq)
    assert sol.retcode == ["Success"]
    from scipy.integrate import solve_ivp
    ref = solve_ivp(lambda t, y: ro.dydtreactor(odom, y, t, odom.p), (0.0, 0.05), odom.y0, method="BDF", rtol=1e-10, atol=1e-20,
                    jac=lambda t, y: ro.jacobiany(odom, y, t, odom.p), t_eval=tq)
    assert ref.status == 0
    Ns = ig.Ns
    assert ref.y[Ns, 1] > 800.0                                  # compression heating
    for k in range(len(tq)):
        yr = ref.y[:, k]
        big = np.abs(yr) > 1e-3 * np.abs(yr[:Ns]).max()
        assert np.max(np.abs(sol.u[0, k] - yr)[big] / np.abs(yr)[big]) < 1e-6, tq[k]
        trace = np.abs(yr) > 1e-12 * np.abs(yr[:Ns]).max()         # the radical pool grows exponentially during the induction time:
                                                                    # both integrations' errors are amplified by the same factor
        assert np.max(np.abs(sol.u[0, k] - yr)[trace] / np.abs(yr)[trace]) < 1e-4, tq[k]
    # ensemble: compression ratios 6..12 through per-reactor profile rows
    n = 32
    cr = np.linspace(6.0, 12.0, n)
    tabs = np.stack([rb.spline_table(ts, V0 * (1.0 - (1.0 - shape) * (1.0 - 1.0 / c) / (1.0 - 0.125))) for c in cr])
    ens = rb.Domain(gdom.kind, ig, np.repeat(gdom.y0, n, axis=0), np.repeat(gdom.cond, n, axis=0))
    ens.profile = tabs
    ens.p, ens._p_default = gdom.p, gdom._p_default
    s = rb.solve(rb.Reactor(ens, None, (0.0, 0.06), p=p).ode, abstol=1e-16, reltol=1e-6)
    assert np.all(s.status == 0)
    Tend = s.yend[:, Ns]
    assert Tend[-1] > Tend[0] + 500.0 and Tend[0] < 1000.0        # the highest compression ratio ignites, the lowest only heats up
    k8 = int(np.argmin(np.abs(cr - 8.0)))
    one = rb.Domain(gdom.kind, ig, gdom.y0, gdom.cond)
    one.profile = tabs[k8]
    one.p, one._p_default = gdom.p, gdom._p_default
    s1 = rb.solve(rb.Reactor(one, None, (0.0, 0.06), p=p).ode, abstol=1e-16, reltol=1e-6)
    assert np.array_equal(s1.yend[0], s.yend[k8])


def test_parametrized_tp_parity_and_trajectory():
    """ParametrizedTPDomain on the reference's test inputs (src/TestReactors.jl:405-423): dydt / jacobiany against the oracle, the
    reference's analytic-vs-AD criterion, and the 0.101 s trajectory against SciPy BDF."""
    ig, oig = _phases()
    ic = {"T": [1000.0, 1400.0, 1500.0], "ts": [0.0, 100.0, 200.0], "P": [1.0e5, 1.8e5, 2.0e5], "H2": 0.67, "O2": 0.33}
    gdom, y0, p = rb.ParametrizedTPDomain(phase=ig, initialconds=ic)
    odom = ro.ParametrizedTPDomain(oig, ic)
    np.testing.assert_allclose(y0, odom.y0, rtol=1e-14)
    ys = cases.states("superminimal_TP", odom.y0, ig.Ns, "ConstantTP")[:3]
    ts = np.array([0.1, 50.0, 180.0])
    par = cases.parameter_vector("superminimal_V", ig.Ns, ig.Nr)
    ens = rb.Domain(gdom.kind, ig, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.profile = gdom.profile
    ens.p, ens._p_default = gdom.p, gdom._p_default
    for pp in (p, par):
        f, J = ens.dydt(ys, pp, t=ts), ens.jacobiany(ys, pp, t=ts)
        for k, y in enumerate(ys):
            fo, M = fl.dydt_terms(odom, y, ts[k], pp)
            rep = fl.report(f[k], fo, fl.EPS * (M + np.abs(fo)), 1e-12, slack=4.0)
            assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, (k, rep)
            Jo, Jf = fl.jac_floor(lambda yy, qq: ro.jacobiany(odom, yy, ts[k], qq), odom, y, pp, t=ts[k])
            rep = fl.report(J[k], Jo, Jf, 1e-12)
            assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, (k, rep["worst_at"], rep["worst_excess"], rep["max_rel_above"])
            Jc = ro.jacobiany_complexstep(odom, y, ts[k], pp)
            assert not np.any(np.abs(J[k] - Jc) > 1e-4 * np.abs(Jc) + 1e-16 * np.abs(Jc).max())
    tq = np.array([0.05, 0.101])
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.101), p=p).ode, abstol=1e-20, reltol=1e-10, saveat=tq)
    assert sol.retcode == ["Success"]
    from scipy.integrate import solve_ivp
    ref = solve_ivp(lambda t, y: ro.dydtreactor(odom, y, t, odom.p), (0.0, 0.101), odom.y0, method="BDF", rtol=1e-11, atol=1e-22,
                    jac=lambda t, y: ro.jacobiany(odom, y, t, odom.p), t_eval=tq)
    assert ref.status == 0
    for k in range(len(tq)):
        yr = ref.y[:, k]
        big = np.abs(yr) > 1e-6 * np.abs(yr[:ig.Ns]).max()
        assert np.max(np.abs(sol.u[0, k] - yr)[big] / np.abs(yr)[big]) < 1e-6


PTV_GOLDEN_BAR = 1e-6       # measured: GPU rtol 1e-9 vs 1e-11 1.4e-8, GPU rtol 1e-11 vs the SciPy golden 2.3e-8
IC_TV = {"ts": [0.0, 600.0, 1200.0], "T": [450.0, 490.0, 500.0], "V": 1.0, "octane": 6154.0, "oxygen": 4.953}      # src/TestReactors.jl:150


def test_parametrized_t_constant_v_parity_and_trajectory():
    """ParametrizedTConstantVDomain on the reference's test inputs (src/TestReactors.jl:148-167): the liquid octane
    mechanism with a temperature ramp at constant volume -- thermo, viscosity, diffusivities and diffusion-limited rate coefficients
    on every call.  dydt / jacobiany against the oracle inside and beyond the profile's knots, the reference's analytic-vs-AD
    criterion, non-default parameters p = [dG; f], the ramp trajectory against SciPy BDF, per-reactor ramps."""
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    d = rb.readinput(cases.fixture("liquid_phase"))
    liq = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase")
    od = ro.readinput(cases.fixture("liquid_phase"))
    oliq = ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0])
    gdom, y0, p = rb.ParametrizedTConstantVDomain(phase=liq, initialconds=IC_TV, constantspecies=["oxygen"])
    odom = ro.ParametrizedTConstantVDomain(oliq, IC_TV, constantspecies=("oxygen",))
    np.testing.assert_array_equal(y0, odom.y0)
    Ns, Nr = liq.Ns, liq.Nr
    ys = cases.states("liquid_TV", odom.y0, Ns, "ConstantTV")[:3]
    ts = np.array([0.0, 700.0, 32977.61568])                    # a knot, inside the ramp, the reference's test time (past the last knot)
    rng = np.random.default_rng(5)
    par = np.concatenate([rng.uniform(-2000.0, 2000.0, Ns), rng.uniform(0.5, 2.0, Nr)])
    ens = rb.Domain(gdom.kind, liq, ys, np.repeat(gdom.cond, len(ys), axis=0), constantspecies=["oxygen"])
    ens.profile = gdom.profile
    ens.p, ens._p_default = gdom.p, gdom._p_default
    for pp in (p, par):
        f, J = ens.dydt(ys, pp, t=ts), ens.jacobiany(ys, pp, t=ts)
        for k, y in enumerate(ys):
            fo, M = fl.dydt_terms(odom, y, ts[k], pp)
            rep = fl.report(f[k], fo, fl.EPS * (M + np.abs(fo)), 1e-12, slack=4.0)
            assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, (k, rep)
            Jo, Jf = fl.jac_floor(lambda yy, qq: ro.jacobiany(odom, yy, ts[k], qq), odom, y, pp, t=ts[k])
            rep = fl.report(J[k], Jo, Jf, 1e-12)
            assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, (k, rep["worst_at"], rep["worst_excess"], rep["max_rel_above"])
            Jc = ro.jacobiany_complexstep(odom, y, ts[k], pp)
            assert not np.any(np.abs(J[k] - Jc) > 1e-4 * np.abs(Jc) + 1e-16 * np.abs(Jc).max())
    # the ramp matters: 450 K at t = 0, 490 K at t = 600
    plain = rb.Domain(gdom.kind, liq, ys, np.repeat(gdom.cond, len(ys), axis=0), constantspecies=["oxygen"])
    plain.profile = gdom.profile
    plain.p, plain._p_default = gdom.p, gdom._p_default
    assert not np.allclose(plain.dydt(ys, p, t=np.zeros(3))[1], ens.dydt(ys, p, t=ts)[1], rtol=1e-2)
    # trajectory through the ramp and one hour beyond it, against the committed SciPy-BDF golden (tests/golden/make_ptv_golden.py:
    # 37 000 evaluations of the NumPy oracle, too slow to repeat here).  As for C3 (tests/test_gpu_large.py) the GPU run is also held
    # to its own tighter run, which separates its error from the golden's.
    gold = np.load(cases.ROOT / "tests" / "golden" / "ptv_liquid.npz")
    tq, yq = gold["tq"], gold["yq"]

    def dist(a, b):
        out = 0.0
        for k in range(len(tq)):
            big = np.abs(b[k]) > 1e-9 * np.abs(b[k]).max()
            out = max(out, float(np.max(np.abs(a[k] - b[k])[big] / np.abs(b[k])[big])))
        return out

    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, float(tq[-1])), p=p).ode, abstol=1e-20, reltol=1e-9, saveat=tq)
    sol11 = rb.solve(rb.Reactor(gdom, y0, (0.0, float(tq[-1])), p=p).ode, abstol=1e-20, reltol=1e-11, saveat=tq)
    assert sol.retcode == ["Success"] and sol11.retcode == ["Success"]
    print("PTV trajectory: GPU rtol 1e-9 vs 1e-11 %.3e, GPU rtol 1e-11 vs SciPy golden %.3e" % (dist(sol.u[0], sol11.u[0]), dist(sol11.u[0], yq)))
    assert dist(sol.u[0], sol11.u[0]) < 1e-6
    assert dist(sol11.u[0], yq) < PTV_GOLDEN_BAR
    # against the same charge held at 450 K, the ramp converts more fuel
    cdom, cy0, cp = rb.ConstantTVDomain(phase=liq, initialconds={"T": 450.0, "V": 1.0, "octane": 6154.0, "oxygen": 4.953}, constantspecies=["oxygen"])
    cold = rb.solve(rb.Reactor(cdom, cy0, (0.0, 4800.0), p=cp).ode, abstol=1e-18, reltol=1e-9)
    io = liq.names.index("octane")
    assert y0[io] - sol.u[0, -1, io] > 2.0 * (y0[io] - cold.yend[0, io])
    # an ensemble of ramps through per-reactor profile rows: reactor k of the batch == the same ramp alone
    n = 16
    tabs = np.stack([rb.spline_table(IC_TV["ts"], [450.0, 450.0 + dT, 450.0 + 1.25 * dT]) for dT in np.linspace(0.0, 60.0, n)])
    ens = rb.Domain(gdom.kind, liq, np.repeat(gdom.y0, n, axis=0), np.repeat(gdom.cond, n, axis=0), constantspecies=["oxygen"])
    ens.profile = tabs
    ens.p, ens._p_default = gdom.p, gdom._p_default
    s = rb.solve(rb.Reactor(ens, None, (0.0, 2400.0), p=p).ode, abstol=1e-16, reltol=1e-6)
    assert np.all(s.status == 0)
    conv = y0[io] - s.yend[:, io]
    assert np.all(np.diff(conv) > 0.0)                          # hotter ramps convert more
    one = rb.Domain(gdom.kind, liq, gdom.y0, gdom.cond, constantspecies=["oxygen"])
    one.profile = tabs[5]
    one.p, one._p_default = gdom.p, gdom._p_default
    s1 = rb.solve(rb.Reactor(one, None, (0.0, 2400.0), p=p).ode, abstol=1e-16, reltol=1e-6)
    assert np.array_equal(s1.yend[0], s.yend[5])


# ---------------------------------------------------------------------------------------------------------------------------------
# jacobianp! and forward sensitivities of the Parametrized* domains.  The reference checks jacobianp against the AD derivative for all
# four (src/TestReactors.jl:163-165,423-425,444-446,465-467): whatever their layout they are variable-k domains, p = [dG; f]
# (src/Domain.jl:3204-3233).  Goldens: tests/golden/sens_param_vectors.npz (make_golden_sens_param.py).
# ---------------------------------------------------------------------------------------------------------------------------------
def _param_domains():
    ig, oig = _phases()
    d = rb.readinput(cases.fixture("liquid_phase"))
    liq = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase")
    ic_tp = {"T": [1000.0, 1400.0, 1500.0], "ts": [0.0, 100.0, 200.0], "P": [1.0e5, 1.8e5, 2.0e5], "H2": 0.67, "O2": 0.33}
    return {"TP": rb.ParametrizedTPDomain(phase=ig, initialconds=ic_tp), "V": rb.ParametrizedVDomain(phase=ig, initialconds=IC_V),
            "P": rb.ParametrizedPDomain(phase=ig, initialconds=IC_P),
            "TV": rb.ParametrizedTConstantVDomain(phase=liq, initialconds=IC_TV, constantspecies=["oxygen"])}


@pytest.mark.parametrize("which", ["TP", "V", "P", "TV"])
def test_parametrized_jacobianp_parity(which):
    gold = np.load(cases.ROOT / "tests" / "golden" / "sens_param_vectors.npz")
    gdom, y0, p = _param_domains()[which]
    ys, ts, Jpo = gold[f"{which}/y"], gold[f"{which}/t"], gold[f"{which}/Jp"]
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, len(ys), axis=0), constantspecies=["oxygen"] if which == "TV" else [])
    ens.profile = gdom.profile
    ens.p, ens._p_default = gdom.p, gdom._p_default
    Jp = ens.jacobianp(ys, p, t=ts)
    # 1e-12 (north_star); the liquid case recomputes 380 diffusion-limited rate coefficients at T(t) through exp(G/RT) with |G/RT| of a few
    # hundred (conditioning ~ 300 eps per rate, the same weight oracle/floors.py gives these terms): 1.02e-12 of the column scale measured
    tol = 1e-12 if which != "TV" else 3e-12
    for k in range(len(ys)):
        assert cases.err_jac(Jp[k], Jpo[k]) < tol, (which, k, cases.err_jac(Jp[k], Jpo[k]))
    assert np.all(Jp[0][Jpo[0] == 0.0] == 0.0)                    # structural zeros stay exact zeros


@pytest.mark.parametrize("which,tend", [("TP", 0.101), ("V", 0.02), ("P", 0.02)])
def test_parametrized_forward_sensitivities_vs_oracle(which, tend):
    from test_gpu_sens import sens_err
    gold = np.load(cases.ROOT / "tests" / "golden" / "sens_param_vectors.npz")
    gdom, y0, p = _param_domains()[which]
    tq, yq, Sq = gold[f"{which}/tq"], gold[f"{which}/yq"], gold[f"{which}/Sq"]
    # induction periods: the sensitivities grow exponentially with the radical pool, and so do both integrations' errors -- golden at
    # rtol 1e-10, GPU at 1e-9 (at 1e-8 / 1e-8 the ParametrizedP case differs by 2e-4 of the column scale)
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p, forwardsensitivities=True).ode, abstol=1e-22, reltol=1e-9, saveat=tq)
    assert sol.retcode[0] == "Success"
    big = np.abs(yq) > 1e-10 * np.max(np.abs(yq), axis=1, keepdims=True)
    assert np.max(np.abs(sol.u[0] - yq)[big] / np.abs(yq)[big]) < 1e-5
    for k in range(len(tq)):
        assert sens_err(sol.sens[0, k], Sq[k]) < 1e-4, (which, k, sens_err(sol.sens[0, k], Sq[k]))
