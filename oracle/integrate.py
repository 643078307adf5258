"""Trajectory oracle (TEST INFRASTRUCTURE): SciPy stiff integrators over the restated RHS.

The reference hands its ODEProblem to SUNDIALS CVODE_BDF (third party, not in /root/reference,
``Sundials = "4"`` in Project.toml:45-76); only converged solution values are pinned by the
reference's tests (src/TestReactors.jl), so any convergent stiff integrator at tight tolerance is a
valid oracle for them.  Integrator-internal behaviour is parity-unpinned.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp

from . import rms_oracle as ro


def solve(domain, tspan, p=None, rtol=1e-6, atol=1e-16, method="BDF", t_eval=None, dense_output=True,
          use_jac=True, **kw):
    """solve(react.ode, CVODE_BDF(); abstol, reltol) stand-in (call sites src/TestReactors.jl:22,197,305)."""
    f = lambda t, y: ro.dydtreactor(domain, y, t, p)
    jac = (lambda t, y: ro.jacobiany(domain, y, t, p)) if use_jac else None
    return solve_ivp(f, tspan, domain.y0, method=method, rtol=rtol, atol=atol, jac=jac, t_eval=t_eval,
                     dense_output=dense_output, **kw)


def idt_grid(sol_fn, Tindex=-1, tmin=1e-15, tmax=0.1, n=10000):
    """The reference's ignition-delay idiom, src/TestReactors.jl:307-308:
    ts = exp.(range(log(1e-15), length=10000, stop=log(0.1))); IDT = ts[argmax(diff(T.(ts)))]."""
    ts = np.exp(np.linspace(np.log(tmin), np.log(tmax), n))
    Ts = np.array([sol_fn(t)[Tindex] for t in ts])
    return ts[np.argmax(np.diff(Ts))]


def idt_continuous(sol, Tindex, domain, p=None):
    """Time of max dT/dt over the accepted steps, refined by a parabola through the three
    neighbouring step points ('IDT definition caveat')."""
    ts = sol.t
    d = np.array([ro.dydtreactor(domain, sol.y[:, i], ts[i], p)[Tindex] for i in range(len(ts))])
    i = int(np.argmax(d))
    if i == 0 or i == len(ts) - 1:
        return ts[i]
    x0, x1, x2 = ts[i - 1], ts[i], ts[i + 1]
    y0, y1, y2 = d[i - 1], d[i], d[i + 1]
    den = (x0 - x1) * (x0 - x2) * (x1 - x2)
    a = (x2 * (y1 - y0) + x1 * (y0 - y2) + x0 * (y2 - y1)) / den
    b = (x2 * x2 * (y0 - y1) + x1 * x1 * (y2 - y0) + x0 * x0 * (y1 - y2)) / den
    return -b / (2 * a) if a < 0 else ts[i]
