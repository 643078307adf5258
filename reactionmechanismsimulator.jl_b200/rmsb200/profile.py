"""Time profiles for the Parametrized* domains (src/Domain.jl:241-498): `getspline(ts, vals)` (src/Tools.jl:46-52) as the table the
device evaluates -- [nt, ts, values, gamma], gamma = second derivatives of the natural cubic spline through the points."""
from __future__ import annotations

import numpy as np


def spline_table(ts, vals) -> np.ndarray:
    """Natural cubic spline through (ts, vals): the lambda -> 0 limit of the smoothing spline the reference fits (lambda = 1e-10)."""
    x = np.asarray(ts, dtype=float)
    g = np.asarray(vals, dtype=float)
    n = x.size
    if n < 2 or g.size != n or np.any(np.diff(x) <= 0):
        raise ValueError("a profile needs at least two increasing times and as many values")
    gam = np.zeros(n)
    if n > 2:
        h = np.diff(x)
        # second derivatives m_1..m_{n-2}: h_{i-1}/6 m_{i-1} + (h_{i-1}+h_i)/3 m_i + h_i/6 m_{i+1} = slope_i - slope_{i-1}
        diag = (h[:-1] + h[1:]) / 3.0
        off = h[1:-1] / 6.0
        slope = np.diff(g) / h
        A = np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)
        gam[1:-1] = np.linalg.solve(A, np.diff(slope))
    return np.concatenate([[float(n)], x, g, gam])


def profile_value(table: np.ndarray, t: float) -> float:
    """Value of the profile at t (host side: initial conditions only; the device has its own evaluator, profile_eval)."""
    n = int(table[0])
    x, g, gam = table[1:1 + n], table[1 + n:1 + 2 * n], table[1 + 2 * n:1 + 3 * n]
    if t <= x[0]:
        h = x[1] - x[0]
        return float(g[0] - (x[0] - t) * ((g[1] - g[0]) / h - h * gam[1] / 6.0))
    if t >= x[-1]:
        h = x[-1] - x[-2]
        return float(g[-1] + (t - x[-1]) * ((g[-1] - g[-2]) / h + h * gam[-2] / 6.0))
    i = int(np.searchsorted(x, t, side="right")) - 1
    h = x[i + 1] - x[i]
    a, b = t - x[i], x[i + 1] - t
    return float((a * g[i + 1] + b * g[i]) / h - a * b * ((1 + a / h) * gam[i + 1] + (1 + b / h) * gam[i]) / 6.0)
