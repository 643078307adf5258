"""Pins the oracle against every golden vector / known-answer test the reference holds for the hot path
(SURVEY.md section 8c), and against the committed oracle fixtures.  CPU only.

Reference sources of the expected values:
  src/TestParse.jl:8-10, src/Calculators/TestRate.jl:6-143, src/Calculators/TestThermo.jl:7-32,
  src/TestTools.jl:5-18, src/TestReactors.jl:193-213 (C1), :301-310 (C2), :29-38 etc. (analytic vs AD Jacobians)
"""
import numpy as np
import pytest

import cases
from oracle import integrate as oi
from oracle import rms_oracle as ro

R, Na = ro.R, ro.Na


def arr(A, n, Ea):
    return {"type": "Arrhenius", "A": A, "n": n, "Ea": Ea}


def isapprox(a, b, rtol):
    """Julia's `a ≈ b rtol=...` for arrays: norm(a-b) <= rtol*max(norm(a), norm(b))."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


# ---------------------------------------------------------------------------------------------
# calculators
# ---------------------------------------------------------------------------------------------
def test_arrhenius_table():  # TestRate.jl:6-9
    a = arr(1e6, 0.5, 41.84 * 1000.0)
    got = [ro.evalkinetics(a, T) for T in np.arange(200.0, 2001.0, 200.0)]
    exp = [1.6721e-4, 6.8770e1, 5.5803e3, 5.2448e4, 2.0632e5, 5.2285e5, 1.0281e6, 1.7225e6, 2.5912e6, 3.6123e6]
    assert isapprox(got, exp, 1e-3)     # the reference compares the vectors with Julia's norm-wise isapprox
    # scalar and vectorised forms agree to rounding (Rate.jl:16 vs Ratevec.jl:24)
    for T in (300.0, 1000.0, 2500.0):
        assert ro.arrhenius_vecform(1e6, 0.5, 41840.0, T) == pytest.approx(ro.arrhenius(a, T), rel=1e-13)


def test_plog_and_multi():  # TestRate.jl:13-44
    a1, a2 = arr(1e6, 1.0, 10e3), arr(1e12, 1.0, 20e3)
    parr = {"type": "PdepArrhenius", "Ps": [0.1e5, 10e5], "arrs": [a1, a2]}
    for T in np.arange(300.0, 1501.0, 100.0):
        assert ro.evalkinetics(parr, T, 1e4) == pytest.approx(ro.arrhenius(a1, T), rel=1e-5)      # clamp low / exact hit
        assert ro.evalkinetics(parr, T, 1e6) == pytest.approx(ro.arrhenius(a2, T), rel=1e-5)      # exact hit high
        assert ro.evalkinetics(parr, T, 1e5) == pytest.approx(np.sqrt(ro.arrhenius(a1, T) * ro.arrhenius(a2, T)), rel=1e-5)
        assert ro.evalkinetics(parr, T, 1e5, vecform=True) == pytest.approx(ro.evalkinetics(parr, T, 1e5), rel=1e-12)
    marr = {"type": "MultiArrhenius", "arrs": [a1, a2]}
    assert ro.evalkinetics(marr, 900.0) == pytest.approx(ro.arrhenius(a1, 900.0) + ro.arrhenius(a2, 900.0), rel=1e-6)


def test_multiplog_table():  # TestRate.jl:48-76
    cm3 = 1e-6
    Ea1, Ea2 = 4740.0 * R, 11200.0 * R
    a = [arr(9.3e-16 * Na * cm3, 0.0, Ea1), arr(9.3e-14 * Na * cm3, 0.0, Ea1), arr(1.4e-11 * Na * cm3, 0.0, Ea2), arr(1.4e-9 * Na * cm3, 0.0, Ea2)]
    Ps = [0.1e5, 10e5]
    mp = {"type": "MultiPdepArrhenius", "parrs": [{"type": "PdepArrhenius", "Ps": Ps, "arrs": a[0:2]},
                                                  {"type": "PdepArrhenius", "Ps": Ps, "arrs": a[2:4]}]}
    Tl = [200.0, 400.0, 600.0, 800.0, 1000.0, 1200.0, 1400.0, 1600.0, 1800.0, 2000.0]
    kexp = np.array([[2.85400e-08, 4.00384e-03, 2.73563e-01, 8.50699e+00, 1.20181e+02, 7.56312e+02, 2.84724e+03, 7.71702e+03, 1.67743e+04, 3.12290e+04],
                     [2.85400e-07, 4.00384e-02, 2.73563e+00, 8.50699e+01, 1.20181e+03, 7.56312e+03, 2.84724e+04, 7.71702e+04, 1.67743e+05, 3.12290e+05],
                     [2.85400e-06, 4.00384e-01, 2.73563e+01, 8.50699e+02, 1.20181e+04, 7.56312e+04, 2.84724e+05, 7.71702e+05, 1.67743e+06, 3.12290e+06]])
    for i, T in enumerate(Tl):
        for j, P in enumerate([1e4, 1e5, 1e6]):
            assert ro.evalkinetics(mp, T, P) == pytest.approx(kexp[j, i], rel=1e-3)
    tb = {"type": "ThirdBody", "arr": a[0], "efficiencies": {4: 1.3}}
    assert ro.evalkinetics(tb, 900.0, C=5.0) == pytest.approx(ro.arrhenius(a[0], 900.0) * 5.0, rel=1e-6)   # TestRate.jl:79-80


HIGH = arr(1.39e16 * 1e-6, -0.534, 2.243e3)
LOW = arr(2.62e33 * 1e-12, -4.76, 10.21e3)


def test_lindemann_troe_tables():  # TestRate.jl:84-120
    lnd = {"type": "Lindemann", "arrhigh": HIGH, "arrlow": LOW, "efficiencies": {1: 1.3}}
    troe = {"type": "Troe", "arrhigh": HIGH, "arrlow": LOW, "a": 0.783, "T3": 74.0, "T1": 2941.0, "T2": 6964.0, "efficiencies": {1: 1.3}}
    Kl = np.array([[1.38023e+08, 2.45661e+08, 2.66439e+08], [6.09146e+07, 2.12349e+08, 2.82604e+08],
                   [4.75671e+06, 4.09594e+07, 1.71441e+08], [7.03616e+05, 6.85062e+06, 5.42111e+07]])
    Kt = np.array([[1.00866e+08, 2.03759e+08, 2.55190e+08], [4.74623e+07, 1.41629e+08, 2.47597e+08],
                   [3.97397e+06, 2.89521e+07, 9.57569e+07], [5.91277e+05, 5.14013e+06, 3.12239e+07]])
    for i, T in enumerate([300.0, 500.0, 1000.0, 1500.0]):
        for j, P in enumerate([1e4, 1e5, 1e6]):
            assert ro.evalkinetics(lnd, T, C=P / (R * T)) == pytest.approx(Kl[i, j], rel=1e-3)
            assert ro.evalkinetics(troe, T, C=P / (R * T)) == pytest.approx(Kt[i, j], rel=1e-3)
            # dk/dCeff closed forms (Rate.jl:232-262) against a complex step
            for kin in (lnd, troe):
                C = P / (R * T)
                d = ro.evalkinetics(kin, T, C=C + 1e-30j).imag / 1e-30
                assert ro.calcdkdCeff(kin, T, C) == pytest.approx(d, rel=1e-9)


def test_chebyshev_table():  # TestRate.jl:124-143
    coefs = np.array([[5.67723e+00, 7.29281e-01, -1.19840e-01, 8.82175e-03], [-1.02669e+00, 8.53639e-01, -3.23485e-02, -2.73670e-02],
                      [-4.47011e-01, 2.44144e-01, 5.59122e-02, -1.01723e-02], [-1.28261e-01, 1.11596e-02, 2.81176e-02, 6.04353e-03],
                      [-1.17034e-02, -2.35646e-02, 6.10090e-04, 4.01309e-03], [1.55433e-02, -1.36846e-02, -4.63048e-03, -2.61353e-04]])
    ch = {"type": "Chebyshev", "coefs": coefs, "Tmin": 300.0, "Tmax": 2000.0, "Pmin": 0.01e5, "Pmax": 100.0e5}
    K = np.array([[2291002.84234755, 1101983.59484935, 43791.93027973, 5201.44365151],
                  [2584518.40955842, 2040371.05643461, 236480.90466183, 41012.26130665],
                  [2572037.53399365, 2574282.31472186, 857727.08665108, 250401.25475447]])
    for i, T in enumerate([300.0, 500.0, 1000.0, 1500.0]):
        for j, P in enumerate([1e4, 1e5, 1e6]):
            assert ro.evalkinetics(ch, T, P) == pytest.approx(K[j, i], rel=1e-3)
            assert ro.chebyshev_vecform(ch, T, P) == pytest.approx(K[j, i], rel=1e-3)


def test_nasa_tables():  # TestThermo.jl:7-32
    polys = [ro.NASAPoly(np.array([4.03055, -0.00214171, 4.90611e-05, -5.99027e-08, 2.38945e-11, -11257.6, 3.5613]), 300.0, 650.73),
             ro.NASAPoly(np.array([-0.307954, 0.0245269, -1.2413e-05, 3.07724e-09, -3.01467e-13, -10693, 22.628]), 650.73, 3000.0)]
    Tl = np.arange(400.0, 2001.0, 200.0)
    Cp = np.array([7.80157, 10.5653, 12.8213, 14.5817, 15.9420, 16.9861, 17.78645, 18.4041, 18.8883]) * R
    S = np.array([29.6534, 33.3516, 36.7131, 39.7715, 42.5557, 45.0952, 47.4179, 49.5501, 51.5152]) * R
    H = np.array([-22.7613, -12.1027, -6.14236, -2.16615, 0.743456, 2.99256, 4.79397, 6.27334, 7.51156]) * Tl * R
    assert isapprox([ro.getHeatCapacity(ro.selectPoly(polys, T), T) for T in Tl], Cp, 1e-3)
    assert isapprox([ro.getEntropy(ro.selectPoly(polys, T), T) for T in Tl], S, 1e-3)
    assert isapprox([ro.getEnthalpy(ro.selectPoly(polys, T), T) for T in Tl], H, 1e-3)
    # the two thermo code paths (0.33333333333 literal vs exact thirds) agree to ~1e-11 but not exactly
    for T in Tl:
        cp, h, s = ro.calcHSCpdless_poly(ro.selectPoly(polys, T).coefs, T)
        G_vec = (h - s) * R * T
        assert G_vec == pytest.approx(ro.getGibbs(ro.selectPoly(polys, T), T), rel=1e-9)


def test_tools():  # TestTools.jl:5-18
    assert ro.evalpoly(2.0, [1.0, 2.0, 3.0]) == 1.0 + 2.0 * 2.0 + 3.0 * 4.0
    x = [1.0, 2.0, 5.0]
    assert ro.getBoundingIndsSorted(0.5, x) == (0, -1)
    assert ro.getBoundingIndsSorted(2.0, x) == (1, -1)
    assert ro.getBoundingIndsSorted(3.0, x) == (1, 2)
    assert ro.getBoundingIndsSorted(7.0, x) == (2, -1)


def test_parse_values():  # TestParse.jl:3-11
    out = ro.readinput(cases.fixture("superminimal"))
    r = out["phase"]["Reactions"][0]
    assert ro.evalkinetics(r.kinetics, 1000.0) == pytest.approx(25618.8, rel=1e-4)
    prod = next(s for s in out["phase"]["Species"] if s.name == r.products[0])
    assert ro.getEnthalpy(ro.selectPoly(prod.polys, 1000.0), 1000.0) == pytest.approx(20677.0, rel=1e-4)
    s5 = out["phase"]["Species"][4]
    assert ro.getEnthalpy(ro.selectPoly(s5.polys, 1000.0), 1000.0) == pytest.approx(20677.2, rel=1e-4)


def test_fixture_structure():
    expect = {"superminimal": (12, 17), "minimal": (26, 71), "ethane": (26, 134), "propane": (82, 1110), "liquid_phase": (35, 380),
              "minimal_edge": (201, 462)}
    for name, (ns, nr) in expect.items():
        d = ro.readinput(cases.fixture(name))
        assert (len(d["phase"]["Species"]), len(d["phase"]["Reactions"])) == (ns, nr)
    d = ro.readinput(cases.fixture("ch4o2cat"))
    assert (len(d["gas"]["Species"]), len(d["surface"]["Species"])) == (12, 29)
    assert (len(d["gas"]["Reactions"]), len(d["surface"]["Reactions"]), len(d[frozenset(("gas", "surface"))]["Reactions"])) == (3, 198, 21)


def test_mcgowan_radius_and_viscosity():
    d = ro.readinput(cases.fixture("liquid_phase"))
    octane = next(s for s in d["phase"]["Species"] if s.name == "octane")
    assert octane.atomnums == {"C": 8, "H": 18} and octane.bondnum == 25
    assert octane.radius == pytest.approx(3.659e-10, rel=1e-3)
    assert d["Solvents"][0].viscosity(450.0) == pytest.approx(1.1709e-4, rel=1e-3)


# ---------------------------------------------------------------------------------------------
# RHS / Jacobian
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", list(cases.CASES))
def test_oracle_matches_committed_vectors(case):
    g = np.load(cases.ROOT / "tests" / "golden" / "parity_vectors.npz")
    dom = cases.oracle_domain(case)
    ys = cases.states(case, dom.y0, dom.phase.Ns, dom.kind)
    np.testing.assert_array_equal(ys, g[f"{case}/y"])
    p = dom.p if dom.kind in ("ConstantV", "ConstantTP") else None
    k = 0 if case.startswith("propane") else slice(None)
    f = np.array([ro.dydtreactor(dom, y, 0.0, p) for y in np.atleast_2d(ys[k])])
    np.testing.assert_allclose(f, np.atleast_2d(g[f"{case}/f"][k]), rtol=1e-13, atol=0)


@pytest.mark.parametrize("case", ["superminimal_V", "superminimal_TP", "ethane_V", "ethane_TP", "liquid_TV", "surface_TAPhi"])
def test_analytic_vs_complexstep_jacobian(case):
    """The reference's own check, src/TestReactors.jl:29-38,219-228,312-321: |Ja - J| <= 1e-4|J| + 1e-16
    (ForwardDiff there, complex step here)."""
    g = np.load(cases.ROOT / "tests" / "golden" / "parity_vectors.npz")
    dom = cases.oracle_domain(case)
    p = dom.p if dom.kind in ("ConstantV", "ConstantTP") else None
    y = g[f"{case}/y"][0]
    Ja = g[f"{case}/J"][0]
    Jc = ro.jacobiany_complexstep(dom, y, 0.0, None if dom.kind != "ConstantV" else p)
    for ind in dom.constantspeciesinds:     # jacobiany! zeroes these rows; the AD Jacobian has them zero through the RHS
        assert np.all(Ja[ind] == 0.0)
    scale = np.max(np.abs(Jc), axis=0, keepdims=True)
    assert np.all(np.abs(Ja - Jc) <= 1e-4 * np.abs(Jc) + 1e-9 * scale + 1e-16)


def test_jacobianp_vs_complexstep():
    """jacobianp vs derivative of the RHS w.r.t. p (src/TestReactors.jl:35-37,225-227,318-320)."""
    for case in ("superminimal_V", "superminimal_TP"):
        dom = cases.oracle_domain(case)
        g = np.load(cases.ROOT / "tests" / "golden" / "parity_vectors.npz")
        y = g[f"{case}/y"][0]
        Jp = ro.jacobianp(dom, y, 0.0, dom.p)
        num = np.zeros_like(Jp)
        for k in range(Jp.shape[1]):
            if dom.kind == "ConstantTP" and k < dom.phase.Ns:
                continue   # dG columns of constant-k domains go through the discontinuous 'thermo changed' branch
            pc = dom.p.astype(complex)
            pc[k] += 1e-30j
            num[:, k] = ro.dydtreactor(dom, y.astype(complex), 0.0, pc).imag / 1e-30
        cols = slice(dom.phase.Ns, None) if dom.kind == "ConstantTP" else slice(None)
        scale = np.max(np.abs(num[:, cols]), axis=0, keepdims=True)
        assert np.all(np.abs(Jp[:, cols] - num[:, cols]) <= 1e-4 * np.abs(num[:, cols]) + 1e-9 * scale + 1e-16)


def test_jacobianp_with_interfaces_reference_criterion():
    """src/TestReactors.jl:277-279: jacobianp(y, p, t, domain, interfaces) of the ConstantTP reactor with an Inlet and an Outlet equals
    the derivative of the RHS (with interfaces) w.r.t. p; and it equals the interface-free jacobianp, as src/Domain.jl:3376-3392 says.
    ConstantV: the Inlet adds the reference's terms to the T and P rows of the dG columns and nothing else (:3421-3430)."""
    dom = cases.oracle_domain("superminimal_TP")
    ifs = [ro.Inlet(dom, {"H2": 0.67, "O2": 0.33, "T": 1000.0, "P": 1e5}, 0.001), ro.Outlet(dom, 0.001)]
    y = np.load(cases.ROOT / "tests" / "golden" / "parity_vectors.npz")["superminimal_TP/y"][0]
    Jp = ro.jacobianp(dom, y, 20.44002454, dom.p, ifs)
    np.testing.assert_array_equal(Jp, ro.jacobianp(dom, y, 20.44002454, dom.p))
    Ns = dom.phase.Ns
    num = np.zeros_like(Jp)
    for k in range(Ns, Jp.shape[1]):                  # rate-coefficient columns (the dG columns of a constant-k domain are not smooth in p)
        pc = dom.p.astype(complex)
        pc[k] += 1e-30j
        num[:, k] = ro.dydtreactor(dom, y.astype(complex), 20.44002454, pc, ifs).imag / 1e-30
    scale = np.max(np.abs(num[:, Ns:]), axis=0, keepdims=True)
    assert np.all(np.abs(Jp[:, Ns:] - num[:, Ns:]) <= 1e-4 * np.abs(num[:, Ns:]) + 1e-9 * scale + 1e-16)
    dv = cases.oracle_domain("superminimal_V")
    iv = [ro.Inlet(dv, {"H2": 0.67, "O2": 0.33, "T": 900.0, "P": 1e5}, 0.002), ro.Outlet(dv, 0.001)]
    yv = np.load(cases.ROOT / "tests" / "golden" / "parity_vectors.npz")["superminimal_V/y"][0]
    J0 = ro.jacobianp(dv, yv, 0.0, dv.p)
    d = ro.jacobianp(dv, yv, 0.0, dv.p, iv) - J0
    Nv = dv.phase.Ns
    assert np.all(d[:Nv] == 0.0) and np.all(d[:, Nv:] == 0.0) and np.all(d[Nv, :Nv] != 0.0)
    th = dv.calcthermo(yv, 0.0, dv.p)
    term = 0.002 * (-yv[:Nv] / th.N) / (th.N * th.Cvave)        # the difference of two Jacobians carries their rounding
    np.testing.assert_allclose(d[Nv, :Nv], term, rtol=1e-9, atol=8 * np.finfo(float).eps * np.abs(J0[Nv, :Nv]).max())
    np.testing.assert_allclose(d[Nv + 1, :Nv], th.P / th.T * term, rtol=1e-9, atol=8 * np.finfo(float).eps * np.abs(J0[Nv + 1, :Nv]).max())


# ---------------------------------------------------------------------------------------------
# trajectories (the reference's pinned reactor results)
# ---------------------------------------------------------------------------------------------
def test_c1_constant_tp_golden():  # src/TestReactors.jl:193-213
    dom = cases.oracle_domain("superminimal_TP")
    sol = oi.solve(dom, (0, 150.11094), p=dom.p, rtol=1e-8, atol=1e-20)
    t = 20.44002454
    y = sol.sol(t)
    N = dom.P * y[-1] / (R * dom.T)
    nm = dom.phase.names
    assert y[nm.index("H2")] / N == pytest.approx(0.412883111, rel=1e-4)
    assert y[nm.index("O2")] / N == pytest.approx(0.200419093, rel=1e-4)
    assert y[nm.index("H2O")] / N == pytest.approx(0.386618602, rel=1e-4)
    th = dom.calcthermo(y)
    rop1 = ro.getrates(dom.phase.rxnarray, th.cs, th.kfs, th.krevs)[0] * th.V     # H+H<=>H2 produces one H2
    assert rop1 == pytest.approx(-5.304207075533127e-13, rel=1e-4)


def test_c2_constant_v_idt_golden():  # src/TestReactors.jl:301-310
    dom = cases.oracle_domain("superminimal_V")
    sol = oi.solve(dom, (0, 0.101), p=dom.p, rtol=1e-8, atol=1e-20)
    assert oi.idt_grid(sol.sol) == pytest.approx(0.038384723436228063, rel=1e-5)
    g = np.load(cases.ROOT / "tests" / "golden" / "parity_vectors.npz")
    assert float(g["superminimal_V/idt_grid"]) == pytest.approx(0.038384723436228063, rel=1e-5)


@pytest.mark.parametrize("case,ifs", [
    ("superminimal_TP", lambda d: [ro.Inlet(d, {"H2": 0.67, "O2": 0.33, "T": 1000.0, "P": 1e5}, lambda t: 0.001), ro.Outlet(d, lambda t: 0.001)]),
    ("superminimal_V", lambda d: [ro.Inlet(d, {"H2": 0.67, "O2": 0.33, "T": 900.0, "P": 1e5}, 0.002), ro.Outlet(d, 0.001)]),
    ("liquid_TV", lambda d: [ro.VolumetricFlowRateInlet(d, {"octane": 6154.0, "oxygen": 4.953, "T": 450.0, "P": 1e5}, 1.0), ro.VolumetricFlowRateOutlet(d, 1.0)]),
])
def test_boundary_interfaces_analytic_jacobian_equals_ad(case, ifs):
    """The reference's own criterion for reactors with interfaces (src/TestReactors.jl:271-280, 53-62): the analytic Jacobian with the
    interface terms equals the derivative of the full RHS, |ja - j| <= 1e-4 |j| + 1e-16 -- here complex step stands in for ForwardDiff."""
    from oracle import rms_oracle as ro2  # noqa: F401
    od = cases.oracle_domain(case)
    y = cases.states(case, od.y0, od.phase.Ns, od.kind)[0]
    po = od.p if od.kind in ("ConstantV", "ConstantTP") else None
    inter = ifs(od)
    ja = ro.jacobiany(od, y, 0.0, po, interfaces=inter)
    j = ro.jacobiany_complexstep(od, y, 0.0, po, interfaces=inter)
    assert not np.any(np.abs(ja - j) > 1e-4 * np.abs(j) + 1e-16 * np.abs(j).max())
    assert np.any(ja != ro.jacobiany(od, y, 0.0, po))


@pytest.mark.parametrize("which", ["V", "P"])
def test_parametrized_domains_analytic_jacobian_equals_ad(which):
    """src/TestReactors.jl:425-467: ParametrizedVDomain / ParametrizedPDomain built from arrays ("ts" + "V" / "P"), analytic Jacobian
    vs the derivative of the full RHS at t = 0.1 (|ja - j| <= 1e-4 |j| + 1e-16); and the profile is the natural cubic spline."""
    from scipy.interpolate import CubicSpline
    d = ro.readinput(cases.fixture("superminimal"))
    ig = ro.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"])
    if which == "V":
        dom = ro.ParametrizedVDomain(ig, {"T": 1000.0, "ts": [0.0, 100.0, 200.0], "V": [0.01, 0.05, 0.1], "H2": 0.67, "O2": 0.33})
        spl, xs, vs = dom.Vf, [0.0, 100.0, 200.0], [0.01, 0.05, 0.1]
    else:
        dom = ro.ParametrizedPDomain(ig, {"T": 1000.0, "ts": [0.0, 100.0, 200.0], "P": [2.0e5, 3.0e5, 5.0e5], "H2": 0.67, "O2": 0.33})
        spl, xs, vs = dom.Pf, [0.0, 100.0, 200.0], [2.0e5, 3.0e5, 5.0e5]
    cs = CubicSpline(xs, vs, bc_type="natural")
    for x in (0.0, 0.1, 63.0, 199.0):
        assert spl(x) == pytest.approx(float(cs(x)), rel=1e-13) and spl.derivative(x) == pytest.approx(float(cs(x, 1)), rel=1e-12)
    y = cases.states("superminimal_V", dom.y0, ig.Ns, "ConstantV")[0]
    ja = ro.jacobiany(dom, y, 0.1, dom.p)
    j = ro.jacobiany_complexstep(dom, y, 0.1, dom.p)
    assert not np.any(np.abs(ja - j) > 1e-4 * np.abs(j) + 1e-16 * np.abs(j).max())


def test_parametrized_t_constant_v_analytic_jacobian_equals_ad():
    """src/TestReactors.jl:148-167: ParametrizedTConstantVDomain on the liquid octane mechanism with "ts" / "T" arrays, analytic
    Jacobian vs the derivative of the full RHS at the reference's test time t = 32977.61568 (past the last knot: the profile
    continues linearly) and inside the ramp; the rate coefficients really follow T(t)."""
    d = ro.readinput(cases.fixture("liquid_phase"))
    liq = ro.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0])
    ic = {"ts": [0.0, 600.0, 1200.0], "T": [450.0, 490.0, 500.0], "V": 1.0, "octane": 6154.0, "oxygen": 4.953}
    dom = ro.ParametrizedTConstantVDomain(liq, ic)
    assert dom.Tf(0.0) == 450.0 and dom.Tf(600.0) == pytest.approx(490.0, rel=1e-14) and dom.Tf(32977.61568) > 500.0
    np.testing.assert_array_equal(dom.p, np.concatenate([np.zeros(liq.Ns), np.ones(liq.Nr)]))
    y = cases.states("liquid_TV", dom.y0, liq.Ns, "ConstantTV")[0]
    for t in (300.0, 32977.61568):
        ja = ro.jacobiany(dom, y, t, dom.p)
        j = ro.jacobiany_complexstep(dom, y, t, dom.p)
        assert not np.any(np.abs(ja - j) > 1e-4 * np.abs(j) + 1e-16)
    # at the first knot the domain is the ConstantTV domain at 450 K (same Gibbs energies to rounding of the two thermo paths)
    ctv = ro.ConstantTVDomain(liq, {"T": 450.0, "V": 1.0, "octane": 6154.0, "oxygen": 4.953})
    np.testing.assert_allclose(ro.dydtreactor(dom, y, 0.0, dom.p), ro.dydtreactor(ctv, y, 0.0, ctv.p), rtol=1e-9, atol=1e-9 * np.abs(ro.dydtreactor(ctv, y, 0.0, ctv.p)).max())
    k450, k490 = dom.calcthermo(y, 0.0).kfs, dom.calcthermo(y, 600.0).kfs
    assert np.max(k490 / k450) > 2.0
