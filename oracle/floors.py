"""Cancellation floors for the elementwise parity tables (TEST INFRASTRUCTURE, see oracle/__init__.py).

The north-star bar is 1e-12 RELATIVE in FP64, but dy/dt and Jacobian entries are sums of signed terms: an entry whose terms
cancel to a fraction 1/kappa of their magnitude cannot be reproduced to better than kappa * eps by ANY FP64 evaluation --
neither by this repo's CUDA path nor by the reference (whose hot path is compiled under @fastmath and re-associates).
This module derives that floor so the tests can report the plain relative error on every entry above it:

* dydt_terms(): the magnitude sum M_i = sum_t kappa_t |term_t| of every component of dydtreactor! -- for the species rows
  V * sum_j |nu_ij| (kappa_f |Rf_j| + kappa_r |Rr_j|) from addreactionratecontributions! (src/Reactor.jl:465-507), for the
  thermodynamic-variable rows the same sums pushed through calcdomainderivatives! (src/Domain.jl:1863-1953).  kappa is the
  conditioning of the exponentials inside a term: k_f = A*exp(x) carries a relative error |x|*eps from the rounding of x, and
  k_rev = k_f/K_c adds sum |nu_i G_i/RT| * eps from the exponent of K_c (50-300 for large molecules) -- any FP64 evaluation,
  the reference's included.  floor_i = eps * M_i.
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
    Gs = th.Gs if len(th.Gs) == Ns else np.asarray(getattr(domain, "Gs", np.zeros(Ns)))     # constant-k domains keep them on the domain
    if p is not None and domain.kind in ("ConstantTP", "ConstantTV", "ConstantTAPhi"):
        Gs = np.asarray(p)[:Ns]
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
        with np.errstate(divide="ignore"):
            kap_f = 2.0 + (abs(np.log(abs(th.kfs[i]))) if th.kfs[i] != 0.0 else 0.0)
        kap_r = kap_f + sum(abs(Gs[ra[j, i]]) for j in range(8) if ra[j, i] >= 0) / (ro.R * th.T)
        mag = kap_f * abs(fR) + kap_r * abs(rR)
        for j in range(8):
            if ra[j, i] >= 0:
                M[ra[j, i]] += mag
    M[:Ns] *= abs(th.V)
    R = ro.R
    if domain.kind in ("ConstantV", "ParametrizedV"):          # src/Domain.jl:1901-1906
        M[Ns] = np.dot(np.abs(th.Us), M[:Ns]) / abs(th.N * th.Cvave)
        M[Ns + 1] = np.sum(M[:Ns]) * R * th.T / th.V + th.P / th.T * M[Ns]
    elif domain.kind in ("ConstantP", "ParametrizedP"):        # :1950-1953
        M[Ns] = np.dot(np.abs(th.Hs), M[:Ns]) / abs(th.N * (th.Cvave + R))
        M[Ns + 1] = np.sum(M[:Ns]) * R * th.T / th.P + M[Ns] * th.V / th.T
    elif domain.kind in ("ConstantTP", "ParametrizedTP"):       # :1863-1866
        M[Ns] = np.sum(M[:Ns]) * R * th.T / th.P
    for i in domain.constantspeciesinds:
        M[i] = 0.0
    return ro.dydtreactor(domain, y, t, p), M


def numeric_floor(fn, inputs, nsample=6, ulps=2.0, seed=0, amplify=16.0):
    """max over `nsample` random perturbations of every entry of every array in `inputs` of |fn(*perturbed) - fn(*inputs)|,
    elementwise, scaled to a relative input uncertainty of `ulps`*eps.  The perturbations are drawn uniformly from
    +-amplify*ulps*eps (continuous magnitudes: with equal magnitudes and random signs the perturbations of two cancelling terms
    coincide exactly in a third of the draws and the floor of such an entry comes out as zero) and the spread is divided by
    `amplify` again -- the response is linear at 1e-15.  None entries of `inputs` are passed through."""
    rng = np.random.default_rng(seed)
    base = np.asarray(fn(*inputs), dtype=float)
    floor = np.zeros_like(base)
    for _ in range(nsample):
        pert = [None if a is None else np.asarray(a, dtype=float) * (1.0 + amplify * ulps * EPS * rng.uniform(-1.0, 1.0, size=np.shape(a))) for a in inputs]
        floor = np.maximum(floor, np.abs(np.asarray(fn(*pert), dtype=float) - base))
    return base, floor / amplify


def closure_floor(domain, y, p, Jabs, t=0.0):
    """Accumulation floor of the thermodynamic-variable ROWS of jacobiany! / jacobianp! (src/Domain.jl:2211-2214,2312-2316,2480-2490,
    3376-3485): they are weighted column sums of the species block (sum_i J_ik * R*T/P, -dot(U, J_:k)/(N*Cv), ...), so their
    rounding floor is eps times the same sums over |J_ik|.  Jabs = |species block| (Ns x ncol); returns the (Ny - Ns) x ncol floor."""
    th = domain.calcthermo(np.asarray(y, dtype=float), t, p)
    Ns = domain.phase.Ns
    R = ro.R
    A = np.abs(Jabs[:Ns])
    if domain.kind in ("ConstantTP", "ParametrizedTP"):
        return EPS * (A.sum(axis=0) * R * th.T / th.P)[None]
    if domain.kind in ("ConstantV", "ParametrizedV"):
        trow = np.abs(th.Us) @ A / abs(th.N * th.Cvave)
        return EPS * np.stack([trow, A.sum(axis=0) * R * th.T / th.V + th.P / th.T * trow])
    if domain.kind in ("ConstantP", "ParametrizedP"):
        trow = np.abs(th.Hs) @ A / abs(th.N * (th.Cvave + R))
        return EPS * np.stack([trow, A.sum(axis=0) * R * th.T / th.P + th.V / th.T * trow])
    return np.zeros((0, A.shape[1]))


def species_block_magnitude(domain, y, p, t=0.0):
    """Sum of the magnitudes of the terms of every entry of the analytic species block of jacobiany! (and of its volume column):
    jacobianynsderiv! (src/Domain.jl:2169-2202) run with every +-deriv accumulated as |deriv|.  An entry like +d - d + small (a
    species on both sides of one reaction next to another reaction's contribution) is only reproducible to eps*d."""
    from unittest import mock
    yv = np.asarray(y, dtype=float)
    th = domain.calcthermo(yv, t, p)
    Ns = domain.phase.Ns
    Vind = Ns if domain.kind in ("ConstantTP", "ParametrizedTP") else (Ns + 1 if domain.kind in ("ConstantP", "ParametrizedP") else None)
    Jabs = np.zeros((len(yv), len(yv)))

    def abs_spread(jac, deriv, slots, col, sign):
        for sl in slots:
            jac[sl, col] += abs(deriv)

    with mock.patch.object(ro, "_spread", abs_spread):
        ro.jacobianynsderiv(Jabs, domain, th, Vind)
    return Jabs


def jacp_magnitude(domain, y, p, t=0.0):
    """The same for the species rows of jacobianp! (src/Domain.jl:3095-3233): |flux terms| of the rate-coefficient columns and the
    accumulated |gderiv| of the dG columns."""
    yv = np.asarray(y, dtype=float)
    ph = domain.phase
    Ns, Nr = ph.Ns, ph.Nr
    th = domain.calcthermo(yv, t, p)
    M = np.zeros((len(yv), Ns + Nr))
    RTinv = 1.0 / (ro.R * th.T)
    ra = ph.rxnarray
    constk = domain.kind in ("ConstantTP", "ConstantTAPhi", "ConstantTV")
    for i in range(Nr):
        rs = [ra[j, i] for j in range(4) if ra[j, i] >= 0]
        ps = [ra[j, i] for j in range(4, 8) if ra[j, i] >= 0]
        prodr = np.prod([th.cs[q] for q in rs])
        prodp = np.prod([th.cs[q] for q in ps])
        if domain.kind == "ConstantTV":
            pk = (domain.p if p is None else p)[Ns + i]
            fderiv = th.kfs[i] * th.kfs[i] / (pk * pk) * prodr
            rderiv = th.kfs[i] * th.krevs[i] / (pk * pk) * prodp
            gderiv = rderiv * (pk * pk) / th.kfs[i] * RTinv
        elif constk:
            if i in ph.efficiencyinds:
                fderiv, rderiv = th.kfs[i] * prodr, th.krevs[i] * prodp
            else:
                fderiv, rderiv = prodr, th.krevs[i] / th.kfs[i] * prodp
            gderiv = rderiv * th.kfs[i] * RTinv
        else:
            fderiv, rderiv = th.kfs[i] * prodr, th.krevs[i] * prodp
            gderiv = rderiv * RTinv
        for q in rs + ps:
            M[q, Ns + i] += abs(fderiv) + abs(rderiv)
        for sp in rs + ps:
            for q in rs + ps:
                M[q, sp] += abs(gderiv)
    M *= abs(th.V)
    return M


def thermo_column_magnitude(domain, y, p, col, t=0.0, h=1e-30):
    """Sum of the magnitudes of the terms of d(dn_i/dt)/d y[col] for a thermodynamic-variable column (T, or P of the ConstantV domain) of the
    variable-k domains: every reaction's forward and reverse rate is differentiated on its own (complex step) and the magnitudes are
    accumulated with the conditioning of the exponentials (see dydt_terms) -- the T / P columns of jacobiany! are the derivative of the
    complete RHS (src/Reactor.jl:1469-1477), so their rounding floor is eps times this sum, not the spread under input perturbations."""
    yc = np.asarray(y, dtype=complex).copy()
    yc[col] += 1j * h
    th = domain.calcthermo(yc, t, p)
    ra = domain.phase.rxnarray
    Ns = domain.phase.Ns
    M = np.zeros(len(y))
    RT = ro.R * th.T.real
    for i in range(ra.shape[1]):
        fR, rR = th.kfs[i], th.krevs[i]
        for j in range(4):
            if ra[j, i] < 0:
                break
            fR = fR * th.cs[ra[j, i]]
        for j in range(4, 8):
            if ra[j, i] < 0:
                break
            rR = rR * th.cs[ra[j, i]]
        kf = abs(th.kfs[i])
        kap_f = 2.0 + (abs(np.log(kf)) if kf != 0.0 else 0.0)
        kap_r = kap_f + sum(abs(th.Gs[ra[j, i]]) for j in range(8) if ra[j, i] >= 0) / RT
        mag = (kap_f * abs(fR.imag) + kap_r * abs(rR.imag)) / h
        for j in range(8):
            if ra[j, i] >= 0:
                M[ra[j, i]] += mag
    M[:Ns] *= abs(th.V.real if hasattr(th.V, "real") else th.V)
    E = th.Us if domain.kind in ("ConstantV", "ParametrizedV") else th.Hs
    den = abs(th.N * th.Cvave) if domain.kind in ("ConstantV", "ParametrizedV") else abs(th.N * (th.Cvave + ro.R))
    M[Ns] = np.dot(np.abs(E), M[:Ns]) / den
    M[Ns + 1] = (np.sum(M[:Ns]) * ro.R * abs(th.T) / abs(th.V) + abs(th.P / th.T) * M[Ns]) if domain.kind in ("ConstantV", "ParametrizedV") else \
        (np.sum(M[:Ns]) * ro.R * abs(th.T) / abs(th.P) + abs(th.V / th.T) * M[Ns])
    return M


def jac_floor(fn, domain, y, p, t=0.0, **kw):
    """(ref, floor) for a Jacobian-like fn(y, p) -> Ny x ncol: conditioning floor (numeric_floor) plus the accumulation floor of the
    closure rows (closure_floor, with a factor for the number of summands' roundings)."""
    mag = kw.pop("magnitude", None)
    ref, floor = numeric_floor(fn, [y, p], **kw)
    Ns = domain.phase.Ns
    if mag is None and ref.shape[1] == ref.shape[0]:
        mag = species_block_magnitude(domain, y, p, t)
    if mag is not None:
        floor[:Ns] += 2.0 * EPS * mag[:Ns]                       # accumulation order of each entry's own terms
    if ref.shape[0] > Ns:
        floor[Ns:] += 4.0 * closure_floor(domain, y, p, np.maximum(np.abs(ref), mag) if mag is not None else ref, t)
    if ref.shape[1] == ref.shape[0] and domain.kind in ("ConstantV", "ConstantP", "ParametrizedV", "ParametrizedP"):
        cols = (Ns, Ns + 1) if domain.kind in ("ConstantV", "ParametrizedV") else (Ns,)
        for col in cols:
            floor[:, col] += EPS * thermo_column_magnitude(domain, y, p, col, t)
    # entries that are EXACTLY zero in the oracle because +d / -d pairs of one reaction cancel (a species on both sides, e.g. a
    # collider): another summation order leaves a residue of a few eps of the column's largest term
    zero = ref == 0.0
    if zero.any():
        col = np.max(np.abs(ref), axis=0, keepdims=True) * np.ones_like(ref)
        floor[zero] += EPS * col[zero]
    return ref, floor


def jacp_floor(domain, y, p, t=0.0, **kw):
    """(ref, floor) of jacobianp!: jac_floor plus the term the adiabatic domains add to the T row of the dG columns,
    -(dot(U, Jp[:, i]) + dn_i/dt) / (N*Cv) (src/Domain.jl:3433-3460): dn_i/dt is itself a cancelling sum whose floor is eps*M_i."""
    ref, floor = jac_floor(lambda yy, pp: ro.jacobianp(domain, yy, t, pp), domain, y, p, t, magnitude=jacp_magnitude(domain, y, p, t), **kw)
    if domain.kind in ("ConstantV", "ConstantP"):
        Ns = domain.phase.Ns
        th = domain.calcthermo(np.asarray(y, dtype=float), t, p)
        _, M = dydt_terms(domain, y, t, p)
        den = abs(th.N * th.Cvave) if domain.kind == "ConstantV" else abs(th.N * (th.Cvave + ro.R))
        add = 4.0 * EPS * M[:Ns] / den
        floor[Ns, :Ns] += add
        floor[Ns + 1, :Ns] += add * (th.P / th.T if domain.kind == "ConstantV" else th.V / th.T)
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
    ratio = err / (bar * np.abs(ref) + slack * floor + 1e-300)
    worst = np.unravel_index(int(np.argmax(ratio)), ratio.shape)
    return {"n": int(x.size), "n_above": int(above.sum()), "max_rel_above": float(rel), "worst_excess": float(ratio[worst]),
            "worst_at": tuple(int(i) for i in worst), "worst_x": float(x[worst]), "worst_ref": float(ref[worst]), "worst_floor": float(floor[worst])}
