"""Forward-sensitivity oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates what `Reactor(...; forwardsensitivities=true)` hands to the integrator (src/Reactor.jl:63-75):
SciMLSensitivity's ODEForwardSensitivityProblem (third party, `SciMLSensitivity = "^7"`, not in /root/reference)
appends S[:, j] = dy/dp_j to the state, u = [y; S[:,1]; ...; S[:,Np]], with
    dS[:, j]/dt = Jy(y, p) S[:, j] + Jp(y, p)[:, j],   S(0) = 0,
Jp = jacobianp! (`paramjac`), and Jy*S the directional derivative of the RHS: the forward-sensitivity ODEFunction is
built WITHOUT `jac` (src/Reactor.jl:65), so SciMLSensitivity differentiates dydt itself; `exact_jac` selects the
complex-step Jacobian of the full RHS as the stand-in for that (default), or the analytic jacobiany! that
threadedsensitivities uses (src/ThreadedSensitivities.jl:59-67); the two agree to rounding.
The reference pins no absolute sensitivity values (src/TestReactors.jl:231-255 only checks forward / adjoint /
threaded against each other at 10 %), so the augmented system integrated here at tight tolerance is the oracle,
cross-checked against central finite differences of the state solution with respect to p.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp

from . import rms_oracle as ro


def augmented_rhs(domain, p, exact_jac=True, interfaces=()):
    Ny = len(domain.y0)
    Np = domain.phase.Ns + domain.phase.Nr

    def f(t, u):
        y = u[:Ny]
        S = u[Ny:].reshape(Np, Ny)                 # S[j] = dy/dp_j
        dy = ro.dydtreactor(domain, y, t, p, interfaces)
        J = ro.jacobiany_complexstep(domain, y, t, p, interfaces=interfaces) if exact_jac else ro.jacobiany(domain, y, t, p, interfaces=interfaces)
        Jp = ro.jacobianp(domain, y, t, p, interfaces)
        dS = S @ J.T + Jp.T
        return np.concatenate([dy, dS.ravel()])
    return f, Ny, Np


def solve_forward(domain, tspan, p, rtol=1e-9, atol=1e-20, t_eval=None, method="BDF", exact_jac=True, interfaces=()):
    """-> (t, y[t][Ny], S[t][Np][Ny]).  `interfaces`: boundary interfaces of the reactor (src/TestReactors.jl:284-296)."""
    f, Ny, Np = augmented_rhs(domain, p, exact_jac, interfaces)
    u0 = np.concatenate([domain.y0, np.zeros(Ny * Np)])
    # Newton matrix of the implicit method: block diagonal, one Jy(y) per block (the sub-diagonal coupling d(Jy*S)/dy is
    # left out -- it only slows the iteration down, the converged solution does not depend on it)
    from scipy.sparse import block_diag

    def jac(t, u):
        J = ro.jacobiany(domain, u[:Ny], t, p, interfaces=interfaces)
        return block_diag([J] * (Np + 1), format="csc")
    sol = solve_ivp(f, tspan, u0, method=method, rtol=rtol, atol=atol, t_eval=t_eval, jac=jac)
    assert sol.success, sol.message
    return sol.t, sol.y[:Ny].T, sol.y[Ny:].T.reshape(-1, Np, Ny)


def finite_difference(domain, tspan, p, j, rel=1e-4, rtol=1e-11, atol=1e-22, t_eval=None, interfaces=()):
    """Central difference of y(t; p) in parameter j (independent of jacobianp)."""
    p = np.asarray(p, dtype=float)
    h = rel * abs(p[j]) if p[j] != 0.0 else rel * ro.R * 1000.0      # dG perturbations: ~ rel * RT
    out = []
    for s in (+1.0, -1.0):
        q = p.copy()
        q[j] += s * h
        f = lambda t, y: ro.dydtreactor(domain, y, t, q, interfaces)
        jac = lambda t, y: ro.jacobiany(domain, y, t, q, interfaces=interfaces)
        sol = solve_ivp(f, tspan, domain.y0, method="BDF", rtol=rtol, atol=atol, jac=jac, t_eval=t_eval)
        assert sol.success
        out.append(sol.y.T)
    return (out[0] - out[1]) / (2.0 * h)
