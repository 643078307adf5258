"""Transitory sensitivities, trapezoidal form (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates `transitorysensitivitiesfulltrapezoidal(sim, t)` (src/TransitorySensitivities.jl:284-302) with its helpers
`getdividingtimescale` (:77-113), `splitnuminds` (:43-58), `getfractionbelow` (:5-15) and
`normalizefulltransitorysensitivities!` (:149-156).  It is a pure consumer of jacobiany / jacobianp: the reference pins
160 of its output values for the ethane ConstantPDomain case (src/TestReactors.jl:373-382), which makes it the reference's
own known-answer test for BOTH Jacobians on a mechanism with falloff and third-body kinetics.  The functions take the two
Jacobians as arrays so the same code checks the oracle's and the CUDA path's."""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm


def getdividingtimescale(Jy, taumax=1e4, taumin=1e-18, taures=10.0 ** 0.5):
    """Time scale that separates the fastest peak of the species time-scale histogram (usediag = true)."""
    with np.errstate(divide="ignore"):
        taus = 1.0 / np.abs(np.diag(Jy))
    edges = 10.0 ** np.arange(np.log10(taumin), np.log10(taumax) + 1e-9, np.log10(taures))
    w = np.zeros(len(edges) - 1, dtype=int)              # StatsBase.fit(Histogram, ...; closed=:left)
    for t in taus:
        if np.isfinite(t) and edges[0] <= t < edges[-1]:
            w[np.searchsorted(edges, t, side="right") - 1] += 1
    N = len(taus)

    def splitnuminds(n):
        c, inds, boo = 0, [], w[0] < n
        for i, x in enumerate(w):
            if boo and x > n:
                c, boo = c + 1, False
                inds.append(i)
            elif (not boo) and x < n:
                c, boo = c + 1, True
                inds.append(i)
        return c, inds

    fracleft = np.cumsum(w) / N
    four, fl, fb = [], [], []
    for n in range(1, int(w.max()) + 1):
        c, inds = splitnuminds(n)
        if c >= 4:
            four.append(inds[1])
            fl.append(fracleft[inds[1]])
            fb.append(sum(min(x, n) for x in w) / N)     # getfractionbelow
    if not four:
        raise ValueError("Could not find a multimodal slice")
    k = int(np.argmin((np.array(fl) - 0.5) ** 2 + (np.array(fb) - 0.5) ** 2))
    return edges[four[k] + 1]


def transitorysensitivitiesfulltrapezoidal(Jy, Jp, y, p, Ns, tau=None):
    """(dSdt, tau): dSdt = (Jp + exp(tau*Jy)*Jp)/2, rate-coefficient columns times p, species rows divided by ns."""
    if tau is None:
        tau = getdividingtimescale(Jy)
    dSdt = (Jp + expm(tau * Jy) @ Jp) / 2.0
    dSdt[:, Ns:] *= np.asarray(p)[Ns:]
    with np.errstate(divide="ignore", invalid="ignore"):      # absent species: 0/0, as in the reference
        dSdt[:Ns, :] /= np.asarray(y)[:Ns, None]
    return dSdt, tau
