"""Multi-domain reactors without interfaces (src/TestReactors.jl:470-531): two ConstantV domains, and a ConstantTP with a ConstantV
domain, integrated as one system reproduce the single-domain solutions in the reference's combined state layout."""
import numpy as np
import pytest

import cases
import rmsb200 as rb
from rmsb200 import _lib as L

pytestmark = pytest.mark.gpu

IC = {"T": 1000.0, "P": 10.0e5, "H2": 0.67, "O2": 0.33}


def _ig():
    d = rb.readinput(cases.fixture("superminimal"))
    return rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")


def test_two_constant_v_domains():
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    ig = _ig()
    Ns = ig.Ns
    domV, y0V, pV = rb.ConstantVDomain(phase=ig, initialconds=IC)
    solV = rb.solve(rb.Reactor(domV, y0V, (0.0, 0.037), p=pV).ode, abstol=1e-16, reltol=1e-6, saveat=[0.03])
    d1, y1, p1 = rb.ConstantVDomain(phase=ig, initialconds=IC)
    d2, y2, p2 = rb.ConstantVDomain(phase=ig, initialconds=IC)
    react, y0, p = rb.Reactor((d1, d2), (y1, y2), (0.0, 0.037), [], (p1, p2))
    assert len(y0) == 2 * (Ns + 2) and np.array_equal(y0[:Ns], y1[:Ns]) and np.array_equal(y0[Ns:2 * Ns], y2[:Ns])
    assert y0[-1] == y1[Ns] and y0[-2] == y1[Ns + 1] and y0[-3] == y2[Ns] and y0[-4] == y2[Ns + 1]       # T1 last, then P1, T2, P2
    assert np.array_equal(p, p1)                                       # one parameter block per distinct phase (src/Reactor.jl:135-165)
    sol = rb.solve(react.ode, abstol=1e-16, reltol=1e-6, saveat=[0.03])
    assert sol.retcode == ["Success"]
    y = sol(0.03)
    ref = solV(0.03)
    np.testing.assert_allclose(y[:Ns], ref[:Ns], rtol=1e-5, atol=1e-30)                  # src/TestReactors.jl:494-495
    np.testing.assert_allclose(y[Ns:-4], ref[:Ns], rtol=1e-5, atol=1e-30)
    ssys = rb.SystemSimulation(sol, (d1, d2), [], p)
    assert ssys.sims[0].molefractions("H2O", 0.03) == pytest.approx(ssys.sims[1].molefractions("H2O", 0.03), rel=1e-12)


def test_constant_tp_with_constant_v_domain():
    ig = _ig()
    Ns = ig.Ns
    dTP, yTP, pTP = rb.ConstantTPDomain(phase=ig, initialconds=IC)
    dV, yV, pV = rb.ConstantVDomain(phase=ig, initialconds=IC)
    solTP = rb.solve(rb.Reactor(dTP, yTP, (0.0, 0.04), p=pTP).ode, abstol=1e-16, reltol=1e-6, saveat=[0.03])
    solV = rb.solve(rb.Reactor(dV, yV, (0.0, 0.04), p=pV).ode, abstol=1e-16, reltol=1e-6, saveat=[0.03])
    react, y0, p = rb.Reactor((dTP, dV), (yTP, yV), (0.0, 0.04), [], (pTP, pV))
    assert len(y0) == (Ns + 1) + (Ns + 2) and y0[-1] == yTP[Ns] and y0[-2] == yV[Ns] and y0[-3] == yV[Ns + 1]
    sol = rb.solve(react.ode, abstol=1e-16, reltol=1e-6, saveat=[0.03])
    y = sol(0.03)
    np.testing.assert_allclose(y[:Ns], solTP(0.03)[:Ns], rtol=1e-5, atol=1e-30)          # src/TestReactors.jl:527-530
    np.testing.assert_allclose(y[Ns:-3], solV(0.03)[:Ns], rtol=1e-5, atol=1e-30)
    with pytest.raises(NotImplementedError):
        rb.Reactor((dTP, dV), (yTP, yV), (0.0, 0.04), [], (pTP, pV), forwardsensitivities=True)
