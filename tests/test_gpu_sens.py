"""Config C5 parity: the CUDA parameter Jacobian (jacobianp!, src/Domain.jl:3095-3485) against the oracle, and the
staggered-direct forward sensitivities (rmsb_sens_batch) against the oracle's augmented ODEForwardSensitivityProblem
system (oracle/sensitivity.py) and against the reference's own C5 check (forward sensitivities of H2 normalised as in
src/TestReactors.jl:231-255).  Tolerances: jacobianp 1e-12 relative (FP64 kernel), sensitivities 1e-4 relative to the
column scale at integrator tolerance 1e-8 (the reference only requires 10 % agreement between its three methods)."""
import numpy as np
import pytest

import cases
import rmsb200 as rb
from rmsb200 import _lib as L

pytestmark = pytest.mark.gpu


class _Golden:
    """tests/golden/sens_vectors.npz (made by tests/golden/make_golden_sens.py), opened on first use."""
    _z = None

    def __getitem__(self, k):
        if _Golden._z is None:
            _Golden._z = np.load(cases.ROOT / "tests" / "golden" / "sens_vectors.npz")
        return _Golden._z[k]

    def __contains__(self, k):
        self["superminimal_V/y"]
        return k in _Golden._z.files


G = _Golden()


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"


def ensemble_of(gdom, ys):
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.constantspeciesinds = gdom.constantspeciesinds
    ens.p, ens._p_default = gdom.p, gdom._p_default
    return ens


def jp_err(Jp, ref, odom, y, p):
    """Mixed relative error of a parameter Jacobian.  Species rows: cases.err_jac.  The ConstantV temperature / pressure rows
    are sums over species, -dot(U, column)/(N*Cv) and sum(column)*R*T/V + P/T*(T row) (src/Domain.jl:3406-3414), that cancel
    by up to 1e6 for thermoneutral reactions: they are judged against the magnitude of the summed terms."""
    Ns = odom.phase.Ns
    if odom.kind != "ConstantV":
        return cases.err_jac(Jp, ref)
    e = cases.err_jac(Jp[:Ns], ref[:Ns])
    th = odom.calcthermo(y, 0.0, p)
    nc = th.N * th.Cvave
    sT = (np.abs(th.Us) @ np.abs(ref[:Ns])) / nc
    sP = np.abs(ref[:Ns]).sum(axis=0) * 8.314 * th.T / th.V + th.P / th.T * sT
    eT = np.max(np.abs(Jp[Ns] - ref[Ns]) / (np.abs(ref[Ns]) + sT + 1e-300))
    eP = np.max(np.abs(Jp[Ns + 1] - ref[Ns + 1]) / (np.abs(ref[Ns + 1]) + sP + 1e-300))
    return max(e, float(eT), float(eP))


@pytest.mark.parametrize("case", list(cases.CASES))
def test_jacobianp_parity(case):
    gdom, y0, p = cases.product_domain(case)
    ys = G[f"{case}/y"]
    ens = ensemble_of(gdom, ys)
    Jp = ens.jacobianp(ys, p)
    odom = cases.oracle_domain(case)
    # 1e-12 (north_star); the dG columns of the 82-species propane mechanism add up several hundred +- contributions per
    # entry in a different association than NumPy's, which leaves up to 1.2e-12 of the column scale
    tol = 1e-12 if odom.phase.Ns < 50 else 3e-12
    for k in range(len(ys)):
        assert jp_err(Jp[k], G[f"{case}/Jp"][k], odom, ys[k], odom.p) < tol
    ref = G[f"{case}/Jp"][0]
    assert np.all(Jp[0][ref == 0.0] == 0.0)                       # structural zeros stay exact zeros
    if f"{case}/p" in G:                                          # non-default parameters (variable-k domain)
        Jpp = ens.jacobianp(ys, G[f"{case}/p"])
        for k in range(len(ys)):
            assert jp_err(Jpp[k], G[f"{case}/Jp_p"][k], odom, ys[k], G[f"{case}/p"]) < tol


def sens_err(S, Sref):
    """max over parameters of |S - Sref| / max_i |Sref[:, par]| (column scale; columns 1e-25 of the largest and below -- the
    oracle's integration noise on the parameters of absent species, exact zeros on the GPU -- are measured against that floor)."""
    sc = np.maximum(np.max(np.abs(Sref), axis=-1, keepdims=True), 1e-25 * np.max(np.abs(Sref))) + 1e-300
    return float(np.max(np.abs(S - Sref) / sc))


@pytest.mark.parametrize("case,tend", [("superminimal_TP", 150.11094), ("superminimal_V", 0.03)])
def test_forward_sensitivities_vs_oracle(case, tend):
    gdom, y0, p = cases.product_domain(case)
    tq = G[f"{case}/tq"]
    react = rb.Reactor(gdom, y0, (0.0, tend), p=p, forwardsensitivities=True)
    sol = rb.solve(react.ode, abstol=1e-20, reltol=1e-8, saveat=tq)
    assert sol.retcode[0] == "Success"
    yq, Sq = G[f"{case}/yq"], G[f"{case}/Sq"]
    big = np.abs(yq) > 1e-10 * np.max(np.abs(yq), axis=1, keepdims=True)
    assert np.max(np.abs(sol.u[0] - yq)[big] / np.abs(yq)[big]) < 1e-5
    # C5 against its golden (SciPy, rtol 1e-9): 1.2e-5 / 9.8e-7 / 1.9e-7 / 4.5e-8 / 7.9e-9 of the column scale at rtol 1e-6 ... 1e-10
    # (scripts/gpu_sens_convergence.py); the ConstantV golden is itself an rtol 1e-8 run through the induction period
    tol = 1e-6 if case == "superminimal_TP" else 1e-4
    for k in range(len(tq)):
        assert sens_err(sol.sens[0, k], Sq[k]) < tol, (k, sens_err(sol.sens[0, k], Sq[k]))
    x, dp = sol.extract_local_sensitivities(float(tq[-1]))
    assert np.allclose(np.array(dp), sol.sens[0, -1])
    xe, dpe = sol.extract_local_sensitivities()
    assert sens_err(np.array(dpe), Sq[-1]) < tol
    # the state computed alongside the sensitivities is the plain solve's state
    plain = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p).ode, abstol=1e-20, reltol=1e-8, saveat=tq)
    assert np.max(np.abs(plain.u[0] - sol.u[0])[big] / np.abs(yq)[big]) < 1e-5


def test_c5_reference_check_normalised_h2_sensitivities():
    """src/TestReactors.jl:231-255: d ln(n_H2) / d ln(k_j) and d ln(n_H2)/dG_i at t = 150.11094 at the reference's own
    tolerances (abstol 1e-21, reltol 1e-7), here against the oracle instead of the adjoint solution; bar 1e-1 there."""
    gdom, y0, p = cases.product_domain("superminimal_TP")
    react = rb.Reactor(gdom, y0, (0.0, 150.11094), p=p, forwardsensitivities=True)
    sol = rb.solve(react.ode, abstol=1e-21, reltol=1e-7)
    x, dp = sol.extract_local_sensitivities()
    Ns = gdom.phase.Ns
    ind = gdom.phase.names.index("H2")
    dpvs = np.array([v[ind] for v in dp])
    dpvs[Ns:] *= np.asarray(p)[Ns:]
    dpvs /= x[ind]
    ref = G["superminimal_TP/Sq"][-1][:, ind].copy()
    ref[Ns:] *= np.asarray(p)[Ns:]
    ref /= G["superminimal_TP/yq"][-1][ind]
    sig = np.abs(ref) > 1e-6 * np.max(np.abs(ref))
    assert np.max(np.abs(dpvs - ref)[sig] / np.abs(ref)[sig]) < 1e-3


def test_sensitivity_ensemble_matches_single_and_finite_difference():
    """An ensemble of ConstantTP conditions: every reactor's sensitivities equal the one-reactor call, and a central
    finite difference of the (GPU) state solution in one rate coefficient reproduces the matching column."""
    gdom, y0, p = cases.product_domain("superminimal_TP")
    ph = gdom.phase
    n = 37
    rng = np.random.default_rng(5)
    T = rng.uniform(950.0, 1150.0, n)
    P = rng.uniform(0.5e5, 3e5, n)
    ns = np.zeros((n, ph.Ns))
    ns[:, ph.names.index("H2")] = rng.uniform(0.4, 0.7, n)
    ns[:, ph.names.index("O2")] = rng.uniform(0.2, 0.4, n)
    ens = rb.ConstantTPEnsemble(ph, T, P, ns)
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 1.0), forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-8)
    assert all(rc == "Success" for rc in sol.retcode)
    assert np.isfinite(sol.sens_end).all()
    for r in (0, 17, 36):
        ic = {"T": T[r], "P": P[r], "H2": ns[r, ph.names.index("H2")], "O2": ns[r, ph.names.index("O2")]}
        d1, y1, p1 = rb.ConstantTPDomain(phase=ph, initialconds=ic)
        s1 = rb.solve(rb.Reactor(d1, y1, (0.0, 1.0), p=p1, forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-8)
        assert np.array_equal(s1.sens_end[0], sol.sens_end[r])            # bitwise: same kernel, same reactor
        j = ph.Ns + 2
        fd = []
        for sgn in (1.0, -1.0):
            q = np.array(p1, dtype=float)
            q[j] *= 1.0 + sgn * 1e-4
            fd.append(rb.solve(rb.Reactor(d1, y1, (0.0, 1.0), p=q).ode, abstol=1e-22, reltol=1e-10).yend[0])
        fd = (fd[0] - fd[1]) / (2e-4 * p1[j])
        assert sens_err(sol.sens_end[r][j][None], fd[None]) < 2e-3


def test_constant_v_sensitivities_through_ignition():
    """ConstantV sensitivities of T to one rate coefficient across the ignition front (t_ign = 0.0384 s) against central
    finite differences of the GPU state solution: they grow by five decades up to ignition and must return to ~0 at
    equilibrium (rate coefficients do not move the equilibrium state)."""
    gdom, y0, p = cases.product_domain("superminimal_V")
    Ns = gdom.phase.Ns
    tq = np.array([0.01, 0.03, 0.037, 0.038, 0.0383, 0.0384, 0.0385, 0.04, 0.05])
    j = Ns + 4
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.05), p=p, forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-9, saveat=tq)
    assert sol.retcode == ["Success"]
    fd = []
    for sgn in (1.0, -1.0):
        q = np.array(p, dtype=float)
        q[j] += sgn * 1e-5
        fd.append(rb.solve(rb.Reactor(gdom, y0, (0.0, 0.05), p=q).ode, abstol=1e-22, reltol=1e-11, saveat=tq).u[0])
    fd = (fd[0] - fd[1]) / 2e-5
    ST = sol.sens[0, :, j, Ns]
    peak = np.abs(ST).max()
    assert peak > 10.0                                            # K per unit multiplier, just before ignition
    assert np.all(np.abs(ST - fd[:, Ns]) <= 0.03 * np.abs(fd[:, Ns]) + 2e-4 * peak)
    assert abs(ST[-1]) < 1e-4 * peak


def test_threadedsensitivities_entry_point():
    """threadedsensitivities(react; odekwargs, senskwargs) of the reference's C5 check (src/TestReactors.jl:234-236) gives the
    forward-sensitivity result (the reference compares the two at 10 %), and `paramindices` selects columns."""
    gdom, y0, p = cases.product_domain("superminimal_TP")
    react = rb.Reactor(gdom, y0, (0.0, 150.11094), p=p)
    th = rb.threadedsensitivities(react, odekwargs={"abstol": 1e-21, "reltol": 1e-7}, senskwargs={"abstol": 1e-21, "reltol": 1e-7})
    fw = rb.solve(rb.Reactor(gdom, y0, (0.0, 150.11094), p=p, forwardsensitivities=True).ode, abstol=1e-21, reltol=1e-7)
    assert np.array_equal(th.sens_end, fw.sens_end)
    sub = rb.threadedsensitivities(react, paramindices=[3, 20], odekwargs={"abstol": 1e-21, "reltol": 1e-7})
    assert sub.sens_end.shape == (1, 2, gdom.Ny) and np.array_equal(sub.sens_end[0, 1], fw.sens_end[0, 19])


@pytest.mark.parametrize("case,tend", [("liquid_TV", 1.0e4), ("surface_TAPhi", 1.0e-6), ("ethane_V", 1.0e-3)])
def test_forward_sensitivities_wider_systems_vs_finite_differences(case, tend):
    """Sensitivities on the wider kernels and the other domains: liquid ConstantTV (35 states, two rows per lane, 415
    parameters, diffusion-limit chain rule, a constant species), surface ConstantTAPhi, ethane ConstantV (falloff kinetics):
    central finite differences of the GPU state solution in the three most influential parameters."""
    gdom, y0, p = cases.product_domain(case)
    Ns = gdom.phase.Ns
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p, forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-9)
    assert sol.retcode == ["Success"] and np.isfinite(sol.sens_end).all()
    S = sol.sens_end[0]
    pv = np.asarray(p, dtype=float)
    infl = np.abs(S).max(axis=1) * np.where(pv != 0.0, np.abs(pv), 8.314 * 1000.0)     # |dy/dln p| (dG columns: per RT)
    picks = [int(np.argmax(infl[Ns:])) + Ns, int(np.argsort(infl[Ns:])[-2]) + Ns, int(np.argmax(infl[:Ns]))]
    for j in picks:
        h = 1e-4 * abs(pv[j]) if pv[j] != 0.0 else 1e-4 * 8.314 * 1000.0
        fd = []
        for sgn in (1.0, -1.0):
            q = pv.copy()
            q[j] += sgn * h
            fd.append(rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=q).ode, abstol=1e-22, reltol=1e-11).yend[0])
        fd = (fd[0] - fd[1]) / (2.0 * h)
        assert np.max(np.abs(S[j] - fd)) <= 5e-3 * np.max(np.abs(fd)) + 1e-30, (case, j)


def test_sensitivity_launch_geometry_small_batches():
    """Small batches with every warps_per_block: the per-warp sensitivity history is indexed by resident warp, so the work
    buffer must cover the launch geometry actually used (rmsb_sens_work_bytes sizes it for the worst case over 0..4), and the
    results never depend on which warp integrates a reactor."""
    gdom, y0, p = cases.product_domain("superminimal_TP")
    for n in (1, 5):
        ens = rb.Domain(gdom.kind, gdom.phase, np.repeat(gdom.y0, n, axis=0), np.repeat(gdom.cond, n, axis=0))
        ens.p, ens._p_default = gdom.p, gdom._p_default
        ref = None
        for wpb in (0, 1, 2, 3, 4):
            sol = rb.solve(rb.Reactor(ens, None, (0.0, 1.0), p=p, forwardsensitivities=True).ode, rb.RMSB200(warps_per_block=wpb),
                           abstol=1e-20, reltol=1e-8)
            assert all(rc == "Success" for rc in sol.retcode) and np.isfinite(sol.sens_end).all()
            if ref is None:
                ref = sol
            assert np.array_equal(sol.sens_end, ref.sens_end) and np.array_equal(sol.yend, ref.yend)
        assert np.array_equal(ref.sens_end[0], ref.sens_end[-1])


# ---------------------------------------------------------------------------------------------------------------------------------
# Parameter Jacobian and forward sensitivities WITH boundary interfaces (src/TestReactors.jl:259-296: the reference runs both on the
# ConstantTP reactor with an Inlet and an Outlet).  Oracle: oracle/sensitivity.py with the interfaces in the RHS, jacobiany and
# jacobianp; goldens tests/golden/sens_if_vectors.npz (make_golden_sens_if.py, finite-difference spot check 3e-9).
# ---------------------------------------------------------------------------------------------------------------------------------
GIF = np.load(cases.ROOT / "tests" / "golden" / "sens_if_vectors.npz")
IF_SETS = {"superminimal_TP": ([("Inlet", ({"H2": 0.67, "O2": 0.33, "T": 1000.0, "P": 1e5}, 0.001)), ("Outlet", (0.001,))], 150.11094),
           "superminimal_V": ([("Inlet", ({"H2": 0.67, "O2": 0.33, "T": 900.0, "P": 1e5}, 0.002)), ("Outlet", (0.001,))], 0.02)}


@pytest.mark.parametrize("case", list(IF_SETS))
def test_jacobianp_and_sensitivities_with_interfaces(case):
    spec, tend = IF_SETS[case]
    gdom, y0, p = cases.product_domain(case)
    gifs = [getattr(rb, nm)(gdom, *args) for nm, args in spec]
    ys, Jpo = GIF[f"{case}/y"], GIF[f"{case}/Jp"]
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.p, ens._p_default = gdom.p, gdom._p_default
    row = rb.pack_interfaces(gdom, gifs)
    Jp = ens.jacobianp(ys, p, interfaces=row)
    for k in range(len(ys)):
        assert cases.err_jac(Jp[k], Jpo[k]) < 1e-12, (case, k)
    plain = ens.jacobianp(ys, p)
    if case == "superminimal_TP":
        assert np.array_equal(Jp, plain)                          # src/Domain.jl:3376-3392: no interface terms
    else:
        Ns = gdom.phase.Ns
        assert np.array_equal(Jp[:, :Ns], plain[:, :Ns]) and not np.array_equal(Jp[:, Ns:, :Ns], plain[:, Ns:, :Ns])
    # forward sensitivities through the reference's call sequence: Reactor(domain, y0, tspan, interfaces; p, forwardsensitivities=true)
    tq, yq, Sq = GIF[f"{case}/tq"], GIF[f"{case}/yq"], GIF[f"{case}/Sq"]
    react = rb.Reactor(gdom, y0, (0.0, tend), gifs, p=p, forwardsensitivities=True)
    sol = rb.solve(react.ode, abstol=1e-20, reltol=1e-8, saveat=tq)
    assert sol.retcode[0] == "Success"
    big = np.abs(yq) > 1e-10 * np.max(np.abs(yq), axis=1, keepdims=True)
    assert np.max(np.abs(sol.u[0] - yq)[big] / np.abs(yq)[big]) < 1e-5
    for k in range(len(tq)):
        assert sens_err(sol.sens[0, k], Sq[k]) < 1e-4, (case, k, sens_err(sol.sens[0, k], Sq[k]))
    # the interfaces matter: the closed reactor's sensitivities are different
    closed = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p, forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-8, saveat=tq)
    assert sens_err(closed.sens[0, -1], Sq[-1]) > 1e-3            # ten times the bar above (0.4 % over the 20 ms of the ConstantV case)
    # an ensemble of residence times through per-reactor interface rows: row k alone == row k of the batch
    n = 8
    rows = np.stack([rb.pack_interfaces(gdom, [getattr(rb, nm)(gdom, *(args[:-1] + (args[-1] * f,))) for nm, args in spec]) for f in np.linspace(0.5, 2.0, n)])
    ensr = rb.Domain(gdom.kind, gdom.phase, np.repeat(gdom.y0, n, axis=0), np.repeat(gdom.cond, n, axis=0))
    ensr.p, ensr._p_default = gdom.p, gdom._p_default
    sb = rb.solve(rb.Reactor(ensr, None, (0.0, tend), rows, p=p, forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-8)
    assert np.all(sb.status == 0)
    one = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), rows[3], p=p, forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-8)
    assert np.array_equal(one.sens_end[0], sb.sens_end[3]) and np.array_equal(one.yend[0], sb.yend[3])
