"""ConstantPDomain (adiabatic reactor at constant pressure, SURVEY section 8 row f1 'other gas domains') on the GPU:
dydt / jacobiany / jacobianp against the oracle, and the reference's own golden for this domain -- the ignition delay of
the ethane mechanism (falloff + third-body kinetics), IDT = 0.07324954954380769 s (src/TestReactors.jl:344-361)."""
import numpy as np
import pytest

import cases
import rmsb200 as rb
from oracle import rms_oracle as ro
from rmsb200 import _lib as L
from test_gpu_sens import jp_err

pytestmark = pytest.mark.gpu
G = np.load(cases.ROOT / "tests" / "golden" / "constp_vectors.npz")
IC = {"T": 1000.0, "P": 2.0e5, "ethane": 1.0, "Ar": 1.0, "O2": 3.5}


@pytest.fixture(scope="module")
def dom():
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    d = rb.readinput(cases.fixture("ethane"))
    ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
    return rb.ConstantPDomain(phase=ig, initialconds=IC)


@pytest.fixture(scope="module")
def odom():
    d = ro.readinput(cases.fixture("ethane"))
    return ro.ConstantPDomain(ro.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"]), IC)


def test_constp_dydt_jacobians_parity(dom, odom):
    gdom, y0, p = dom
    np.testing.assert_allclose(y0, odom.y0, rtol=1e-14)
    ys = G["y"]
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.p, ens._p_default = gdom.p, gdom._p_default
    f, J, Jp = ens.dydt(ys, p), ens.jacobiany(ys, p), ens.jacobianp(ys, p)
    fp, Jpp = ens.dydt(ys, G["p"]), ens.jacobianp(ys, G["p"])
    for k in range(len(ys)):
        assert cases.err_dydt(f[k], G["f"][k]) < 1e-12
        assert cases.err_jac(J[k], G["J"][k]) < 1e-12
        assert cases.err_dydt(fp[k], G["f_p"][k]) < 1e-12
        assert jp_err_p(Jp[k], G["Jp"][k], odom, ys[k], odom.p) < 1e-12
        assert jp_err_p(Jpp[k], G["Jp_p"][k], odom, ys[k], G["p"]) < 1e-12


def jp_err_p(Jp, ref, odom, y, p):
    """As test_gpu_sens.jp_err for the ConstantP thermo rows: -dot(H, column)/(N*Cp) and sum(column)*R*T/P + V/T*(T row)
    cancel for thermoneutral reactions, so they are judged against the magnitude of the summed terms."""
    Ns = odom.phase.Ns
    e = cases.err_jac(Jp[:Ns], ref[:Ns])
    th = odom.calcthermo(y, 0.0, p)
    nc = th.N * (th.Cvave + 8.314)
    sT = (np.abs(th.Hs) @ np.abs(ref[:Ns])) / nc
    sV = np.abs(ref[:Ns]).sum(axis=0) * 8.314 * th.T / th.P + th.V / th.T * sT
    eT = np.max(np.abs(Jp[Ns] - ref[Ns]) / (np.abs(ref[Ns]) + sT + 1e-300))
    eV = np.max(np.abs(Jp[Ns + 1] - ref[Ns + 1]) / (np.abs(ref[Ns + 1]) + sV + 1e-300))
    return max(e, float(eT), float(eV))


def test_constp_reference_ignition_delay_golden(dom):
    """src/TestReactors.jl:354-361: solve with abstol 1e-16 / reltol 1e-6, IDT on the 10 000-point log grid up to 0.2 s."""
    gdom, y0, p = dom
    ts = np.exp(np.linspace(np.log(1e-15), np.log(0.2), 10000))
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 1.0), p=p), abstol=1e-16, reltol=1e-6, saveat=ts)
    assert sol.retcode == ["Success"]
    idt = rb.idt_grid(sol, tmax=0.2)
    assert idt == pytest.approx(0.07324954954380769, rel=1e-5)
    assert sol.idt[0] == pytest.approx(idt, rel=4e-3)              # continuous definition, within a grid cell
    sim = rb.Simulation(sol, gdom)
    assert sim.getP(0.05) == 2.0e5 and sim.getT(0.2) > 3000.0


def test_constp_trajectory_and_ensemble(dom, odom):
    gdom, y0, p = dom
    tq, yq = G["tq"], G["yq"]
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 1.0), p=p), abstol=1e-20, reltol=1e-10, saveat=tq)
    big = np.abs(yq) > 1e-12 * np.max(np.abs(yq[:, :gdom.phase.Ns]))
    assert np.max(np.abs(sol.u[0] - yq)[big] / np.abs(yq)[big]) < 2e-6
    # an ensemble over T0 / P: every reactor ignites, pressure stays put, enthalpy-conserving final temperatures are ordered
    n = 64
    rng = np.random.default_rng(9)
    T0, P = rng.uniform(950.0, 1200.0, n), 10 ** rng.uniform(5.0, 6.0, n)
    ns = np.zeros((n, gdom.phase.Ns))
    for name, v in (("ethane", 1.0), ("Ar", 1.0), ("O2", 3.5)):
        ns[:, gdom.phase.names.index(name)] = v
    ens = rb.ConstantPEnsemble(gdom.phase, T0, P, ns)
    s2 = rb.solve(rb.Reactor(ens, None, (0.0, 10.0)), abstol=1e-16, reltol=1e-6)
    assert np.all(s2.status == 0)
    Ns = gdom.phase.Ns
    assert np.all(s2.yend[:, Ns] > T0 + 1500.0)
    # ideal gas at constant pressure: V_end = N_end*R*T_end/P
    np.testing.assert_allclose(s2.yend[:, Ns + 1], s2.yend[:, :Ns].sum(axis=1) * 8.314 * s2.yend[:, Ns] / P, rtol=1e-5)


def test_constp_reference_transitory_sensitivity_golden(dom):
    """The reference's 160 transitory-sensitivity goldens (src/TestReactors.jl:373-382) recomputed from the CUDA path's
    state y(0.01), jacobiany and jacobianp (the matrix exponential and the normalisation are host post-processing, as in the
    reference)."""
    from oracle import transitory as tr
    gold = np.loadtxt(cases.ROOT / "tests" / "golden" / "transitory_ethane_reference.txt")
    gdom, y0, p = dom
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.01), p=p), abstol=1e-16, reltol=1e-6, saveat=[0.01])
    # the product's consumer (rmsb200/transitory.py through the Simulation accessor, as the reference calls it) ...
    dSdt, tau = rb.Simulation(sol, gdom).transitorysensitivitiesfulltrapezoidal(0.01)
    # ... and the oracle's restatement fed with the same CUDA Jacobians agree to rounding
    y = sol.yend[0]
    one = rb.Domain(gdom.kind, gdom.phase, y[None], gdom.cond)
    one.p, one._p_default = gdom.p, gdom._p_default
    dS_o, tau_o = tr.transitorysensitivitiesfulltrapezoidal(one.jacobiany(y[None], p)[0], one.jacobianp(y[None], p)[0], y, p, gdom.phase.Ns)
    assert tau == tau_o
    fin = np.isfinite(dS_o)
    np.testing.assert_allclose(dSdt[fin], dS_o[fin], rtol=1e-9, atol=1e-300)
    row = dSdt[gdom.phase.names.index("ethane"), :]
    ok = np.abs(row - gold) <= 1e-4 * np.abs(gold) + 1e-16
    assert ok.sum() >= 159 and np.all(np.abs(row - gold) <= 2e-4 * np.abs(gold) + 1e-16)


def test_constp_forward_sensitivities_vs_finite_differences(dom):
    """Forward sensitivities of the adiabatic constant-pressure reactor (28 states, 160 parameters, R = 1 kernel with five
    parameter chunks) in the induction period, against central finite differences of the GPU state solution."""
    gdom, y0, p = dom
    Ns = gdom.phase.Ns
    tend = 0.05
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p, forwardsensitivities=True).ode, abstol=1e-20, reltol=1e-9)
    assert sol.retcode == ["Success"] and np.isfinite(sol.sens_end).all()
    S = sol.sens_end[0]
    norms = np.abs(S[:, Ns])                                   # sensitivity of T
    for j in (Ns + int(np.argmax(norms[Ns:])), int(np.argmax(norms[:Ns])), Ns + int(np.argsort(norms[Ns:])[-3])):
        pj = float(np.asarray(p)[j])
        h = 1e-4 * abs(pj) if pj != 0.0 else 1e-4 * 8.314 * 1000.0
        fd = []
        for sgn in (1.0, -1.0):
            q = np.array(p, dtype=float)
            q[j] += sgn * h
            fd.append(rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=q).ode, abstol=1e-22, reltol=1e-11).yend[0])
        fd = (fd[0] - fd[1]) / (2.0 * h)
        assert np.max(np.abs(S[j] - fd)) <= 2e-3 * np.max(np.abs(fd)) + 1e-30
