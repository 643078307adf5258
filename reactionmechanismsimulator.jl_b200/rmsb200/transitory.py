"""Transitory sensitivities (src/TransitorySensitivities.jl): a consumer of the batched Jacobians.

`transitorysensitivitiesfulltrapezoidal(sim, t)` in the reference costs one `jacobiany!`, one `jacobianp!` and one matrix
exponential per (reactor, time) (src/TransitorySensitivities.jl:284-302).  Here the two Jacobians of EVERY requested state come
from one `rmsb_jac_batch` and one `rmsb_jacp_batch` call; the time-scale histogram, exp(tau*Jy)*Jp and the normalisation are
host post-processing exactly as in the reference (it is not part of the integration hot path).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def _histogram_left_closed(values: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """StatsBase.fit(Histogram, values, edges; closed=:left) -- counts per bin, values outside the edges dropped."""
    v = values[np.isfinite(values)]
    v = v[(v >= edges[0]) & (v < edges[-1])]
    return np.bincount(np.searchsorted(edges, v, side="right") - 1, minlength=len(edges) - 1)[:len(edges) - 1]


def _crossings(weights: np.ndarray, level: int):
    """splitnuminds (src/TransitorySensitivities.jl:43-58): bin indices at which the histogram crosses `level`
    (upwards crossings need weight > level, downwards ones weight < level)."""
    below = weights[0] < level
    idx = []
    for i, w in enumerate(weights):
        if below and w > level:
            below = False
            idx.append(i)
        elif not below and w < level:
            below = True
            idx.append(i)
    return idx


def getdividingtimescale(Jy: np.ndarray, taumax: float = 1e4, taumin: float = 1e-18, taures: float = 10.0 ** 0.5) -> float:
    """getdividingtimescale(Jy; taumax, taumin, taures, usediag=true) (src/TransitorySensitivities.jl:77-113): the edge of the
    time-scale histogram (tau_i = 1/|Jy_ii|) that separates its fastest mode, chosen among all horizontal slices that cut the
    histogram at least four times as the one nearest to (half the species left of it, half the histogram mass below the slice)."""
    with np.errstate(divide="ignore"):
        taus = 1.0 / np.abs(np.diagonal(Jy))
    nb = int(np.floor((np.log10(taumax) - np.log10(taumin)) / np.log10(taures) + 1e-9))
    edges = 10.0 ** (np.log10(taumin) + np.log10(taures) * np.arange(nb + 1))
    w = _histogram_left_closed(taus, edges)
    n_tot = taus.size
    left = np.cumsum(w) / n_tot
    best, best_cost = None, np.inf
    for level in range(1, int(w.max()) + 1):
        idx = _crossings(w, level)
        if len(idx) < 4:
            continue
        second = idx[1]
        frac_below = np.minimum(w, level).sum() / n_tot          # getfractionbelow (:5-15)
        cost = (left[second] - 0.5) ** 2 + (frac_below - 0.5) ** 2
        if cost < best_cost:
            best, best_cost = second, cost
    if best is None:
        raise ValueError("Could not find a multimodal slice of the time-scale histogram")
    return float(edges[best + 1])


def transitorysensitivitiesfulltrapezoidal(domain, y, p=None, tau=float("nan"), normalized: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Batched transitorysensitivitiesfulltrapezoidal (src/TransitorySensitivities.jl:284-302) for the states y [n][Ny] of the
    ensemble `domain`: dSdt = (Jp + exp(tau*Jy)*Jp)/2 per state; tau = NaN picks every state's own dividing time scale, tau = 0
    returns Jp.  normalized: rate-coefficient columns times p, species rows divided by the mole numbers
    (normalizefulltransitorysensitivities!, :149-156).  Returns (dSdt [n][Ny][Ns+Nr], tau [n])."""
    from scipy.linalg import expm
    y = np.ascontiguousarray(np.atleast_2d(y), dtype=float)
    n = y.shape[0]
    Ns = domain.phase.Ns
    pv = domain.p if p is None else p
    Jp = domain.jacobianp(y, p)
    taus = np.broadcast_to(np.asarray(tau, dtype=float), (n,)).copy()
    if np.all(taus == 0.0):
        out = Jp.copy()
    else:
        Jy = domain.jacobiany(y, p)
        out = np.empty_like(Jp)
        for r in range(n):
            if np.isnan(taus[r]):
                taus[r] = getdividingtimescale(Jy[r])
            out[r] = Jp[r] if taus[r] == 0.0 else 0.5 * (Jp[r] + expm(taus[r] * Jy[r]) @ Jp[r])
    if normalized:
        pr = np.atleast_2d(np.asarray(pv, dtype=float))
        out[:, :, Ns:] *= pr[:, None, Ns:]
        with np.errstate(divide="ignore", invalid="ignore"):     # absent species: 0/0 as in the reference
            out[:, :Ns, :] /= y[:, :Ns, None]
    return out, taus
