"""Simulation accessors (src/Simulation.jl:13-39,146-254,357-373,763-778): pure host code over the
arrays `solve` returned, plus batched `rmsb_rates_batch` calls for rates / ROPs.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .mechanism import R
from .reactor import Solution


class Simulation:
    """Simulation(sol, domain) for reactor `r` of an ensemble solution."""

    def __init__(self, sol: Solution, domain=None, r: int = 0):
        self.sol, self.domain, self.r = sol, (domain or sol.reactor.domain), r
        self.names = list(self.domain.phase.names)
        self.species = self.domain.phase.species
        self.reactions = self.domain.phase.reactions
        self.p = sol.reactor.p

    def spcindex(self, name: str) -> int:
        return self.names.index(name)          # 0-based (reference: 1-based)

    def y(self, t):
        return self.sol(t, self.r)

    def _TPV(self, y):
        d, Ns = self.domain, self.domain.phase.Ns
        c = d.cond[self.r]
        if d.kind == L.DOMAIN_CONSTANT_V:
            return y[Ns], y[Ns + 1], c[0]
        if d.kind == L.DOMAIN_CONSTANT_P:
            return y[Ns], c[0], y[Ns + 1]
        if d.kind == L.DOMAIN_CONSTANT_TP:
            return c[0], c[1], y[Ns]
        if d.kind == L.DOMAIN_CONSTANT_TV:
            return c[0], 1.0e8, c[1]
        return c[0], 0.0, c[1]

    def getT(self, t):
        return self._TPV(self.y(t))[0]

    def getP(self, t):
        return self._TPV(self.y(t))[1]

    def getV(self, t):
        return self._TPV(self.y(t))[2]

    def N(self, t):
        """Total moles: P*V/(R*T) for gas domains, sum(ns) otherwise (src/Domain.jl:919,1068,1516)."""
        y = self.y(t)
        T, P, V = self._TPV(y)
        if self.domain.kind in (L.DOMAIN_CONSTANT_V, L.DOMAIN_CONSTANT_TP, L.DOMAIN_CONSTANT_P):
            return P * V / (R * T)
        return float(np.sum(y[:self.domain.phase.Ns]))

    def getC(self, t):
        return self.N(t) / self.getV(t)

    def molefractions(self, name=None, t=None):
        """molefractions(sim[, name], t)  src/Simulation.jl:149-161"""
        y = self.y(t)
        x = y[:self.domain.phase.Ns] / self.N(t)
        return x if name is None else x[self.spcindex(name)]

    def concentrations(self, name=None, t=None):
        """concentrations(sim[, name], t)  src/Simulation.jl:165-204"""
        y = self.y(t)
        c = y[:self.domain.phase.Ns] / self.getV(t)
        return c if name is None else c[self.spcindex(name)]

    def rates(self, t):
        """rates(sim, t): getrate(rxn, cs, kfs, krevs) * V for every reaction, mol/s (src/Simulation.jl:763-767)."""
        d = self.domain
        y = self.y(t)[None, :]
        sub = _single(d, self.r)
        kf, kr = sub.rates(y, self.p)
        V = self._TPV(y[0])[2]
        cs = y[0, :d.phase.Ns] / V
        ra = d.phase.rxnarray
        out = np.zeros(d.phase.Nr)
        for j in range(d.phase.Nr):
            f, b = kf[0, j], kr[0, j]
            for k in range(4):
                if ra[j, k] > 0:
                    f *= cs[ra[j, k] - 1]
                if ra[j, 4 + k] > 0:
                    b *= cs[ra[j, 4 + k] - 1]
            out[j] = (f - b) * V
        return out

    def transitorysensitivitiesfulltrapezoidal(self, t, tau=float("nan"), normalized=True):
        """transitorysensitivitiesfulltrapezoidal(sim, t; tau, normalized) -> (dSdt, tau)  src/TransitorySensitivities.jl:284-302"""
        from .transitory import transitorysensitivitiesfulltrapezoidal as tsens
        dS, taus = tsens(_single(self.domain, self.r), self.y(t)[None, :], self.p, tau, normalized)
        return dS[0], float(taus[0])

    def rops(self, name, t):
        """rops(sim, name, t): (#product - #reactant occurrences) * rate * V per reaction (src/Simulation.jl:357-373)."""
        d = self.domain
        rts = self.rates(t)
        s = self.spcindex(name) + 1
        ra = d.phase.rxnarray
        out = np.zeros(d.phase.Nr)
        for j in range(d.phase.Nr):
            nu = int(np.sum(ra[j, 4:] == s)) - int(np.sum(ra[j, :4] == s))
            out[j] = nu * rts[j]
        return out


def _single(dom, r):
    """View of one reactor of an ensemble as a Domain (for per-time accessor calls)."""
    from .domain import Domain
    if dom.n == 1:
        return dom
    sub = Domain(dom.kind, dom.phase, dom.y0[r], dom.cond[r], ())
    sub.constantspeciesinds = dom.constantspeciesinds
    sub.p, sub._p_default = dom.p, dom._p_default
    return sub


def ignition_delay(sol: Solution, min_rise: float = 400.0) -> np.ndarray:
    """Ignition delay per reactor: time of max dT/dt (continuous form of the reference's idiom
    ts[argmax(diff(T.(ts)))], src/TestReactors.jl:307-308); NaN for reactors whose temperature rose less
    than `min_rise` K by t_end (no ignition)."""
    dom = sol.reactor.domain
    assert dom.kind in (L.DOMAIN_CONSTANT_V, L.DOMAIN_CONSTANT_P)
    Ti = dom.phase.Ns
    out = sol.idt.copy()
    out[(sol.yend[:, Ti] - sol.reactor.y0[:, Ti]) < min_rise] = np.nan
    return out


def idt_grid(sol: Solution, r: int = 0, Tindex=None, tmin=1e-15, tmax=0.1, n=10000):
    """The reference's grid definition applied to a saved trajectory (needs `saveat` = that grid)."""
    ts = np.exp(np.linspace(np.log(tmin), np.log(tmax), n))
    Tindex = sol.reactor.domain.phase.Ns if Tindex is None else Tindex
    assert sol.t is not None and sol.t.size == n and np.allclose(sol.t, ts, rtol=1e-14, atol=0)
    return ts[np.argmax(np.diff(sol.u[r, :, Tindex]))]
