"""Oracle for config C4: gas ConstantTPDomain + surface ConstantTAPhiDomain coupled through a
ReactiveInternalInterfaceConstantTPhi (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows src/Reactor.jl:99-258 (multi-domain Reactor: species blocks first in domain order, thermodynamic
variables packed from the end of y), :662-725 (multi-domain dydtreactor!), src/Interface.jl:71-133,170-232
(interface construction, constant k's, sticking-coefficient upgrade) and src/PhaseState.jl:131-162
(two-phase getKc).  The reference's Jacobian for multi-domain systems is ForwardDiff of the whole RHS
(src/Reactor.jl:203); the complex-step Jacobian below is the stand-in.
"""
from __future__ import annotations

import math
from typing import List

import numpy as np

from . import rms_oracle as ro
from .rms_oracle import R, Na, kB


def stickingcoefficient2arrhenius(sc: dict, sitedensity: float, N: int, mw: float) -> dict:
    """src/Interface.jl:220-232.  The reference least-squares-fits log k(T) = log(s(T)/sigma^N*sqrt(kB*T/(2*pi*m))) over
    300:10:2000 K with an Arrhenius form; whenever s(T) = A*T^n*exp(-Ea/RT) < 1 on that range the target IS exactly
    Arrhenius with A' = A/sigma^N*sqrt(kB/(2*pi*m)), n' = n + 1/2, Ea' = Ea, so the fit converges to the closed form
    (all 19 sticking reactions of ch4o2cat.rms satisfy max s(T) = 0.84 < 1)."""
    Ts = np.arange(300.0, 2000.0 + 1e-9, 10.0)
    s = sc["A"] * Ts ** sc["n"] * np.exp(-sc["Ea"] / (R * Ts))
    if np.any(s >= 1.0):
        raise NotImplementedError("sticking coefficient saturates (s >= 1) inside 300-2000 K: needs the LsqFit path")
    mass = mw / Na
    return {"type": "Arrhenius", "A": sc["A"] / sitedensity ** N * math.sqrt(kB / (2.0 * math.pi * mass)), "n": sc["n"] + 0.5, "Ea": sc["Ea"]}


class CoupledGasSurface:
    """Reactor((domaingas, domaincat), (y0gas, y0cat), tspan, (inter,), (pgas, pcat, pinter)) for a ConstantTP gas domain,
    a ConstantTAPhi surface domain and one ReactiveInternalInterfaceConstantTPhi.  y = [ns_gas; ns_surf; V]."""
    kind = "CoupledTPTAPhi"

    def __init__(self, gas: ro.ConstantTPDomain, surf: ro.ConstantTAPhiDomain, interface_reactions: List[ro.Reaction], A: float):
        assert gas.T == surf.T                                   # src/Interface.jl:90
        self.gas, self.surf, self.T, self.P, self.A = gas, surf, gas.T, gas.P, A
        g, s = gas.phase, surf.phase
        self.N1, self.N2 = g.Ns, s.Ns
        self.Ns = g.Ns + s.Ns
        self.Ny = self.Ns + 1
        self.names = g.names + s.names
        self.y0 = np.concatenate([gas.y0[:g.Ns], surf.y0, [gas.y0[g.Ns]]])
        # interface: upgradekinetics + getinterfacereactioninds + constant kfs/krevs (src/Interface.jl:89-108)
        self.irxns = []
        idx = {n: i for i, n in enumerate(self.names)}
        ra = -np.ones((8, len(interface_reactions)), dtype=np.int64)
        kfs, krevs = [], []
        for i, r in enumerate(interface_reactions):
            kin = r.kinetics
            if kin["type"] == "StickingCoefficient":
                gasspc = [n for n in r.reactants if n in g.spcdict]
                assert len(gasspc) == 1
                mw = g.species[g.spcdict[gasspc[0]]].molecularweight
                kin = stickingcoefficient2arrhenius(kin, s.sitedensity, len(r.reactants) - 1, mw)
            self.irxns.append(ro.Reaction(i, r.reactants, r.products, kin, r.reversible, r.forwardable))
            for j, n in enumerate(r.reactants):
                ra[j, i] = idx[n]
            for j, n in enumerate(r.products):
                ra[j + 4, i] = idx[n]
            kf = ro.evalkinetics(kin, self.T)                                                   # getkf(rxn, nothing, T, ...)
            dG, dN1, dN2 = 0.0, 0, 0                                                            # getKc two-phase, PhaseState.jl:131-162
            for n in r.reactants:
                if n in g.spcdict:
                    dG -= gas.Gs[g.spcdict[n]]; dN1 -= 1
                else:
                    dG -= surf.Gs[s.spcdict[n]]; dN2 -= 1
            for n in r.products:
                if n in g.spcdict:
                    dG += gas.Gs[g.spcdict[n]]; dN1 += 1
                else:
                    dG += surf.Gs[s.spcdict[n]]; dN2 += 1
            Kc = math.exp(-dG / (R * self.T)) * g.getC0(self.T) ** dN1 * s.getC0(self.T) ** dN2
            kfs.append(kf * float(r.forwardable))
            krevs.append(kf / Kc * float(r.reversible))
        self.irxnarray, self.ikfs, self.ikrevs = ra, np.array(kfs), np.array(krevs)
        # per-domain rxnarrays re-indexed into the global species numbering (src/Reactor.jl:108-116)
        self.gra = g.rxnarray.copy()
        self.sra = np.where(s.rxnarray >= 0, s.rxnarray + self.N1, -1)

    def dydt(self, y, t=0.0):
        """src/Reactor.jl:662-725."""
        y = np.asarray(y)
        N1, Ns = self.N1, self.Ns
        V = y[Ns]
        dt = np.result_type(y.dtype, float)
        cstot = np.zeros(self.Ny, dtype=dt)
        dydt = np.zeros(self.Ny, dtype=dt)
        cstot[:N1] = y[:N1] / V
        ro.addreactionratecontributions(dydt, self.gra, cstot, self.gas.kfs, self.gas.krevs)
        dydt[:N1] *= V
        cstot[N1:Ns] = y[N1:Ns] / self.A
        ro.addreactionratecontributions(dydt, self.sra, cstot, self.surf.kfs, self.surf.krevs)
        dydt[N1:Ns] *= self.A
        ro.addreactionratecontributions(dydt, self.irxnarray, cstot, self.ikfs, self.ikrevs, self.A)   # evaluate(), Interface.jl:115-117
        dydt[Ns] = np.sum(dydt[:N1]) * R * self.T / self.P                                           # calcdomainderivatives! TP
        return dydt

    def term_scale(self, y):
        """sum_j |nu_ij| (|forward flux_j| + |reverse flux_j|) per component: the magnitude against which rounding of
        dy_i/dt has to be judged (fast adsorption/desorption pairs cancel to a net 1e6 times smaller than their terms)."""
        y = np.asarray(y, dtype=float)
        N1, Ns = self.N1, self.Ns
        cs = np.concatenate([y[:N1] / y[Ns], y[N1:Ns] / self.A])
        sc = np.zeros(self.Ny)
        for ra, kf, kr, fac in ((self.gra, self.gas.kfs, self.gas.krevs, y[Ns]), (self.sra, self.surf.kfs, self.surf.krevs, self.A),
                                (self.irxnarray, self.ikfs, self.ikrevs, self.A)):
            for i in range(ra.shape[1]):
                fR, rR = kf[i], kr[i]
                for j in range(4):
                    if ra[j, i] >= 0:
                        fR *= cs[ra[j, i]]
                    if ra[j + 4, i] >= 0:
                        rR *= cs[ra[j + 4, i]]
                for j in range(8):
                    if ra[j, i] >= 0:
                        sc[ra[j, i]] += (abs(fR) + abs(rR)) * fac
        sc[Ns] = np.sum(sc[:N1]) * R * self.T / self.P
        return sc

    def jacobian(self, y, h=1e-30):
        y = np.asarray(y, dtype=float)
        J = np.zeros((self.Ny, self.Ny))
        for k in range(self.Ny):
            yc = y.astype(complex)
            yc[k] += 1j * h
            J[:, k] = self.dydt(yc).imag / h
        return J


def ch4o2cat_system(fixture_path: str, T=800.0, P=1.0e5, xCH4=0.1):
    """The reference's own C4 set-up, src/TestReactors.jl:534-556 (sitedensity 2.486e-5, A = 1e5*V)."""
    d = ro.readinput(fixture_path)
    ig = ro.IdealGas(d["gas"]["Species"], d["gas"]["Reactions"])
    cat = ro.IdealSurface(d["surface"]["Species"], d["surface"]["Reactions"], 2.486e-5)
    ic = {"T": T, "P": P, "O2": 0.2, "N2": 0.7, "CH4": xCH4}
    gas = ro.ConstantTPDomain(ig, ic)
    V = gas.y0[ig.Ns]
    A = 1.0e5 * V
    surf = ro.ConstantTAPhiDomain(cat, {"T": T, "A": A, "vacantX": cat.sitedensity * A})
    inter = d[frozenset(("gas", "surface"))]["Reactions"]
    return CoupledGasSurface(gas, surf, inter, A)
