"""Integrator robustness on the reference's own tolerance habits (src/TestReactors.jl uses abstol 1e-20 with reltol from
1e-6 to 1e-12): no Newton-failure storms from exact-zero components, no step-size collapse at reltol 1e-12, no
temperature drift in the post-ignition tail.  Each of these failed in the first version of the integrator (see DESIGN.md
section 5, 'robustness')."""
import numpy as np
import pytest

import cases
import rmsb200 as rb
from rmsb200 import _lib as L
from test_gpu_parity import sweep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"


@pytest.mark.parametrize("fs", [False, True])
def test_tolerance_sweep_abstol_1e20(fs):
    """ConstantV ignition at abstol 1e-20 for reltol 1e-6 ... 1e-12, with and without forward sensitivities: converges, the
    step count grows smoothly with the tolerance, collider species that no reaction changes (Ar, N2) stay exactly zero."""
    gdom, y0, p = cases.product_domain("superminimal_V")
    Ns = gdom.phase.Ns
    steps, Tend = [], []
    for rtol in (1e-6, 1e-7, 1e-8, 1e-10, 1e-12):
        sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.05), p=p, forwardsensitivities=fs).ode, abstol=1e-20, reltol=rtol)
        st = sol.stats[0]
        assert sol.retcode == ["Success"], (rtol, sol.retcode)
        assert st["n_newton_fail"] <= 5 and st["n_err_fail"] < 0.2 * st["n_steps"], (rtol, st)
        for name in ("Ar", "He", "Ne", "N2"):
            assert sol.yend[0, gdom.phase.names.index(name)] == 0.0
        steps.append(int(st["n_steps"]))
        Tend.append(sol.yend[0, Ns])
    assert steps == sorted(steps) and steps[0] < 800 and steps[-1] < 8000, steps
    assert np.allclose(Tend, Tend[-1], rtol=1e-6)


def test_no_temperature_drift_in_equilibrium_tail():
    """After ignition the state is at equilibrium: T must stay put from 2x the ignition delay to t_end (a stale Jacobian at
    step sizes 1e4 times larger used to let it coast by 1e-4), and P*V/(R*T) must keep tracking sum(ns)."""
    gdom, _, _ = cases.product_domain("superminimal_V")
    ig = gdom.phase
    T0, P0, ns = sweep(ig, 4000)
    ens = rb.ConstantVEnsemble(ig, T0, P0, ns)
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-16, reltol=1e-6, saveat=[0.05, 0.1])
    assert np.all(sol.status == 0)
    ign = np.isfinite(sol.idt) & (sol.idt < 0.02) & (sol.yend[:, ig.Ns] > T0 + 400)
    assert ign.sum() > 1500
    drift = np.abs(sol.u[ign, 1, ig.Ns] / sol.u[ign, 0, ig.Ns] - 1.0)
    assert drift.max() < 5e-6, drift.max()
    Ni = sol.yend[:, ig.Ns + 1] * ens.cond[:, 0] / (8.314 * sol.yend[:, ig.Ns])
    np.testing.assert_allclose(Ni, sol.yend[:, :ig.Ns].sum(axis=1), rtol=1e-5)
