"""Coupled gas + catalytic-surface systems (config C4): the reference's multi-domain
`Reactor((domaingas, domaincat), (y0gas, y0cat), tspan, (inter,), (pgas, pcat, pinter))`
(src/Reactor.jl:99-258,662-725) with a `ReactiveInternalInterfaceConstantTPhi` (src/Interface.jl:71-133).

The two phases and the interface reactions are merged into ONE two-phase mechanism descriptor (species of the gas
first, then the surface; reactions gas | surface | interface, each block in its own phase order) so that the whole
coupled system is one reactor of domain kind COUPLED_TP_TAPHI on the device: y = [ns_gas; ns_surf; V],
cond = [T, P, A].
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from .domain import Domain
from .mechanism import ElementaryReaction, Na, R, kB
from .phase import Phase


def stickingcoefficient2arrhenius(sc: dict, sitedensity: float, N: int, mw: float) -> dict:
    """src/Interface.jl:220-232.  The reference fits log k(T), k = s(T)/sigma^N*sqrt(kB*T/(2*pi*m)), with an Arrhenius form
    over 300:10:2000 K (LsqFit).  As long as the sticking coefficient s(T) = A*T^n*exp(-Ea/RT) stays below its cap of 1 on
    that range the target is exactly Arrhenius -- A' = A/sigma^N*sqrt(kB/(2*pi*m)), n' = n + 1/2, Ea' = Ea -- and the fit
    returns these values; a saturating s(T) would need the numerical fit and is rejected here."""
    Ts = np.arange(300.0, 2000.0 + 1e-9, 10.0)
    if np.any(sc["A"] * Ts ** sc["n"] * np.exp(-sc["Ea"] / (R * Ts)) >= 1.0):
        raise NotImplementedError("sticking coefficient reaches 1 inside 300-2000 K: outside the accelerated hot path")
    mass = mw / Na
    return {"type": "Arrhenius", "A": sc["A"] / sitedensity ** N * math.sqrt(kB / (2.0 * math.pi * mass)), "n": sc["n"] + 0.5, "Ea": sc["Ea"]}


class CoupledPhase(Phase):
    """Two-phase mechanism: gas phase + surface phase + interface reactions."""

    def __init__(self, gas: Phase, surf: Phase, interface_reactions: Sequence[ElementaryReaction], device: Optional[int] = None):
        assert gas.kind == L.PHASE_GAS and surf.kind == L.PHASE_SURFACE
        self.gas, self.surf = gas, surf
        self.kind, self.name, self.solvent, self.sitedensity = L.PHASE_GAS, f"{gas.name}+{surf.name}", None, surf.sitedensity
        self.diffusionlimited = False
        self.species = list(gas.species) + list(surf.species)
        self.names = [s.name for s in self.species]
        self.spcdict = {n: i + 1 for i, n in enumerate(self.names)}
        # upgradekinetics (src/Interface.jl:196-218): sticking coefficients -> Arrhenius
        irx = []
        for r in interface_reactions:
            kin = r.kinetics
            if kin["type"] == "StickingCoefficient":
                gasspc = [s for s in r.reactants if s.name in gas.spcdict]
                assert len(gasspc) == 1
                if gasspc[0].molecularweight <= 0.0:
                    raise ValueError(f"species {gasspc[0].name} needs a molecular weight for the sticking-coefficient conversion")
                kin = stickingcoefficient2arrhenius(kin, surf.sitedensity, len(r.reactants) - 1, gasspc[0].molecularweight)
            irx.append(ElementaryReaction(0, r.reactants, r.products, kin, r.reversible, r.forwardable, r.electronchange))
        merged = list(gas.reactions) + list(surf.reactions) + irx
        self.reactions = [ElementaryReaction(i + 1, r.reactants, r.products, r.kinetics, r.reversible, r.forwardable, r.electronchange)
                          for i, r in enumerate(merged)]
        self.n_gas_rxn, self.n_surf_rxn, self.n_iface_rxn = gas.Nr, surf.Nr, len(irx)
        self.Ns, self.Nr = len(self.species), len(self.reactions)
        self.nvec = 0
        self.veckineticsinds = []
        self.rxnarray = np.zeros((self.Nr, 8), dtype=np.int32)
        for j, r in enumerate(self.reactions):
            for k, s in enumerate(r.reactants):
                self.rxnarray[j, k] = self.spcdict[s.name]
            for k, s in enumerate(r.products):
                self.rxnarray[j, 4 + k] = self.spcdict[s.name]
        self.reversibility = np.array([r.reversible for r in self.reactions], dtype=np.uint8)
        self.forwardability = np.array([r.forwardable for r in self.reactions], dtype=np.uint8)
        self.efficiencyinds = []
        if gas.efficiencyinds or surf.efficiencyinds:
            raise NotImplementedError("third-body efficiencies in a coupled gas-surface system")
        self._device = gas._device if device is None else device
        # single-phase reactions keep their phase's vectorised / scalar forms; interface reactions use getkf (scalar form)
        self._vectorised = list(gas._vectorised) + list(surf._vectorised) + [False] * len(irx)
        self._species_phase = np.array([0] * gas.Ns + [1] * surf.Ns, dtype=np.int32)
        self._rxn_domain = np.array([0] * gas.Nr + [1] * surf.Nr + [2] * len(irx), dtype=np.int32)
        self._handle = None
        self._build_desc()


class ReactiveInternalInterfaceConstantTPhi:
    """ReactiveInternalInterfaceConstantTPhi(domain1, domain2, reactions, T, A, phi=0.0) -> (inter, pinter)
    (src/Interface.jl:89-108); use `make_interface`."""

    def __init__(self, domain1: Domain, domain2: Domain, reactions, T: float, A: float):
        assert domain1.kind == L.DOMAIN_CONSTANT_TP and domain2.kind == L.DOMAIN_CONSTANT_TAPHI
        assert float(domain1.cond[0, 0]) == float(domain2.cond[0, 0]) == float(T)      # src/Interface.jl:90
        self.domain1, self.domain2, self.T, self.A = domain1, domain2, float(T), float(A)
        self.phase = CoupledPhase(domain1.phase, domain2.phase, reactions)
        self.reactions = self.phase.reactions[-self.phase.n_iface_rxn:] if self.phase.n_iface_rxn else []
        self.domaininds = [0, 1]


def make_interface(domain1, domain2, reactions, T, A, phi=0.0):
    assert phi == 0.0
    inter = ReactiveInternalInterfaceConstantTPhi(domain1, domain2, reactions, T, A)
    dom = coupled_domain(inter)
    kf, _ = dom.rates()
    inter.kfs = kf[0, -inter.phase.n_iface_rxn:].copy()
    inter.p = inter.kfs.copy()
    return inter, inter.p


def coupled_domain(inter: ReactiveInternalInterfaceConstantTPhi) -> Domain:
    """The merged system as one Domain: y0 = [ns_gas; ns_surf; V] (src/Reactor.jl:106-132), cond = [T, P, A]."""
    dg, dc = inter.domain1, inter.domain2
    N1 = dg.phase.Ns
    y0 = np.concatenate([dg.y0[:, :N1], dc.y0, dg.y0[:, N1:N1 + 1]], axis=1)
    cond = np.concatenate([dg.cond, np.full((dg.n, 1), inter.A)], axis=1)
    names = [inter.phase.names[i - 1] for i in list(dg.constantspeciesinds)] + [dc.phase.names[i - 1] for i in list(dc.constantspeciesinds)]
    dom = Domain(L.DOMAIN_COUPLED_TP_TAPHI, inter.phase, y0, cond, names)
    dom.indexes = [1, inter.phase.Ns, inter.phase.Ns + 1]
    dom.thermovariabledict = {"V": inter.phase.Ns + 1}
    return dom


def CoupledEnsemble(inter_phase: CoupledPhase, T, P, A, ns_gas, ns_surf) -> Domain:
    """n coupled gas-surface reactors from arrays: T[n], P[n], A[n], ns_gas[n][N1], ns_surf[n][N2]; V = N*R*T/P."""
    T, P, A = (np.asarray(x, dtype=float) for x in (T, P, A))
    ns_gas, ns_surf = np.atleast_2d(np.asarray(ns_gas, dtype=float)), np.atleast_2d(np.asarray(ns_surf, dtype=float))
    V = ns_gas.sum(axis=1) * R * T / P
    y0 = np.concatenate([ns_gas, ns_surf, V[:, None]], axis=1)
    return Domain(L.DOMAIN_COUPLED_TP_TAPHI, inter_phase, y0, np.stack([T, P, A], axis=1))


class SystemSimulation:
    """SystemSimulation(sol, domains, interfaces, p) (src/Simulation.jl:92-134): accessors over the coupled solution."""

    def __new__(cls, sol, domains=None, interfaces=(), p=None, r: int = 0):
        if type(sol).__name__ == "MultiDomainSolution":          # domains without interfaces: one Simulation per domain
            from .multidomain import MultiDomainSimulation
            return MultiDomainSimulation(sol, domains, interfaces, p)
        return object.__new__(cls)

    def __init__(self, sol, domains, interfaces, p=None, r: int = 0):
        from .simulation import Simulation
        self.sol, self.domains, self.interfaces, self.r = sol, domains, interfaces, r
        self.phase = interfaces[0].phase
        self.names = list(self.phase.names)
        self.N1 = domains[0].phase.Ns
        self.A = float(sol.reactor.domain.cond[r, 2])
        self.T, self.P = float(sol.reactor.domain.cond[r, 0]), float(sol.reactor.domain.cond[r, 1])
        outer = self

        class _Gas:   # ssys.sims[1] of the reference: the ConstantTP gas domain
            names = list(domains[0].phase.names)

            def molefractions(self, name=None, t=None):
                y = outer.sol(t, outer.r)
                N = outer.P * y[outer.phase.Ns] / (R * outer.T)
                x = y[:outer.N1] / N
                return x if name is None else x[self.names.index(name)]

        class _Surf:
            names = list(domains[1].phase.names)

            def concentrations(self, name=None, t=None):
                y = outer.sol(t, outer.r)
                c = y[outer.N1:outer.phase.Ns] / outer.A
                return c if name is None else c[self.names.index(name)]

        self.sims = [_Gas(), _Surf()]

    def concentrations(self, name, t):
        """concentrations(ssys, name, t): n/V for gas species, n/A for surface species."""
        y = self.sol(t, self.r)
        i = self.names.index(name)
        return y[i] / (y[self.phase.Ns] if i < self.N1 else self.A)
