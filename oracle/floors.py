"""Cancellation floors for the elementwise parity tables (TEST INFRASTRUCTURE, see oracle/__init__.py).

The north-star bar is 1e-12 RELATIVE in FP64, but dy/dt and Jacobian entries are sums of signed terms: an entry whose terms
cancel to a fraction 1/kappa of their magnitude cannot be reproduced to better than kappa * eps by ANY FP64 evaluation --
neither by this repo's CUDA path nor by the reference (whose hot path is compiled under @fastmath and re-associates).
This module derives that floor so the tests can report the plain relative error on every entry above it:

* dydt_terms(): the magnitude sum M_i = sum_t |term_t| of every component of dydtreactor! -- for the species rows
  V * sum_j |nu_ij| (|Rf_j| + |Rr_j|) from addreactionratecontributions! (src/Reactor.jl:465-507), for the thermodynamic-variable
  rows the same sums pushed through calcdomainderivatives! (src/Domain.jl:1863-1953).  floor_i = eps * M_i.
* numeric_floor(): for the Jacobians (whose terms are the derivatives of those, src/Reactor.jl:899-1477, Domain.jl:3095-3485)
  the floor is measured instead of re-derived: the spread of the oracle's own result when every input (state, conditions,
  parameters) is perturbed by a random relative 2 ulp.  For an entry sum_t m_t(u) of monomial terms this spread is
  eps * sum_t deg_t |m_t| up to sign correlations, i.e. the same term-magnitude bound.
"""
from __future__ import annotations

import numpy as np

from . import rms_oracle as ro

EPS = 2.220446049250313e-16


def dydt_terms(domain, y, t=0.0, p=None):
    """(f, M): dydtreactor! and the per-component sum of the magnitudes of its terms."""
    y = np.asarray(y, dtype=float)
    th = domain.calcthermo(y, t, p)
    ra = domain.phase.rxnarray
    Ns = domain.phase.Ns
    M = np.zeros(len(y))
    for i in range(ra.shape[1]):
        fR = th.kfs[i]
        for j in range(4):
            if ra[j, i] < 0:
                break
            fR = fR * th.cs[ra[j, i]]
        rR = th.krevs[i]
        for j in range(4, 8):
            if ra[j, i] < 0:
                break
            rR = rR * th.cs[ra[j, i]]
        mag = abs(fR) + abs(rR)
        for j in range(8):
            if ra[j, i] >= 0:
                M[ra[j, i]] += mag
    M[:Ns] *= abs(th.V)
    R = ro.R
    if domain.kind == "ConstantV":          # src/Domain.jl:1901-1906
        M[Ns] = np.dot(np.abs(th.Us), M[:Ns]) / abs(th.N * th.Cvave)
        M[Ns + 1] = np.sum(M[:Ns]) * R * th.T / th.V + th.P / th.T * M[Ns]
    elif domain.kind == "ConstantP":        # :1950-1953
        M[Ns] = np.dot(np.abs(th.Hs), M[:Ns]) / abs(th.N * (th.Cvave + R))
        M[Ns + 1] = np.sum(M[:Ns]) * R * th.T / th.P + M[Ns] * th.V / th.T
    elif domain.kind == "ConstantTP":       # :1863-1866
        M[Ns] = np.sum(M[:Ns]) * R * th.T / th.P
    for i in domain.constantspeciesinds:
        M[i] = 0.0
    return ro.dydtreactor(domain, y, t, p), M


def numeric_floor(fn, inputs, nsample=6, ulps=2.0, seed=0):
    """max over `nsample` random perturbations (relative size `ulps`*eps on every entry of every array in `inputs`) of
    |fn(*perturbed) - fn(*inputs)|, elementwise.  None entries of `inputs` are passed through."""
    rng = np.random.default_rng(seed)
    base = np.asarray(fn(*inputs), dtype=float)
    floor = np.zeros_like(base)
    for _ in range(nsample):
        pert = [None if a is None else np.asarray(a, dtype=float) * (1.0 + ulps * EPS * rng.choice([-1.0, 1.0], size=np.shape(a))) for a in inputs]
        floor = np.maximum(floor, np.abs(np.asarray(fn(*pert), dtype=float) - base))
    return base, floor


def closure_floor(domain, y, p, Jabs, t=0.0):
    """Accumulation floor of the thermodynamic-variable ROWS of jacobiany! / jacobianp! (src/Domain.jl:2211-2214,2312-2316,2480-2490,
    3376-3485): they are weighted column sums of the species block (sum_i J_ik * R*T/P, -dot(U, J_:k)/(N*Cv), ...), so their
    rounding floor is eps times the same sums over |J_ik|.  Jabs = |species block| (Ns x ncol); returns the (Ny - Ns) x ncol floor."""
    th = domain.calcthermo(np.asarray(y, dtype=float), t, p)
    Ns = domain.phase.Ns
    R = ro.R
    A = np.abs(Jabs[:Ns])
    if domain.kind == "ConstantTP":
        return EPS * (A.sum(axis=0) * R * th.T / th.P)[None]
    if domain.kind == "ConstantV":
        trow = np.abs(th.Us) @ A / abs(th.N * th.Cvave)
        return EPS * np.stack([trow, A.sum(axis=0) * R * th.T / th.V + th.P / th.T * trow])
    if domain.kind == "ConstantP":
        trow = np.abs(th.Hs) @ A / abs(th.N * (th.Cvave + R))
        return EPS * np.stack([trow, A.sum(axis=0) * R * th.T / th.P + th.V / th.T * trow])
    return np.zeros((0, A.shape[1]))


def jac_floor(fn, domain, y, p, t=0.0, **kw):
    """(ref, floor) for a Jacobian-like fn(y, p) -> Ny x ncol: conditioning floor (numeric_floor) plus the accumulation floor of the
    closure rows (closure_floor, with a factor for the number of summands' roundings)."""
    ref, floor = numeric_floor(fn, [y, p], **kw)
    Ns = domain.phase.Ns
    if ref.shape[0] > Ns:
        floor[Ns:] += 4.0 * closure_floor(domain, y, p, ref, t)
    return ref, floor


def report(x, ref, floor, bar=1e-12, slack=8.0):
    """Elementwise comparison of x with ref given the absolute floor of every entry.
    Returns dict: n, n_above (entries whose floor is below bar/10 of their magnitude: a relative comparison at `bar` is meaningful
    there), max_rel_above (plain |x-ref|/|ref| over those), worst_excess (max over ALL entries of
    |x-ref| / (bar*|ref| + slack*floor): <= 1 means every entry is within the bar or within `slack` floors)."""
    x, ref, floor = (np.asarray(a, dtype=float) for a in (x, ref, floor))
    err = np.abs(x - ref)
    above = (np.abs(ref) > 0) & (floor <= 0.1 * bar * np.abs(ref))
    rel = np.max(err[above] / np.abs(ref[above])) if above.any() else 0.0
    excess = np.max(err / (bar * np.abs(ref) + slack * floor + 1e-300))
    return {"n": int(x.size), "n_above": int(above.sum()), "max_rel_above": float(rel), "worst_excess": float(excess)}
