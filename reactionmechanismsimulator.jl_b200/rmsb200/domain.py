"""Domains: the reference's ConstantVDomain / ConstantTPDomain / ConstantTVDomain / ConstantTAPhiDomain
constructors (src/Domain.jl:42-102,120-170,500-560,660-711) returning ``(domain, y0, p)``, plus array
constructors that build a whole ensemble of conditions at once (the reference parses one Dict per
reactor; an ensemble of 100k conditions is built here from arrays without a Python loop).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import ctypes as C

import numpy as np

from . import _lib as L
from .mechanism import R
from .phase import Phase


class Domain:
    """One condition (n = 1) or an ensemble (n > 1) of reactors of the same kind sharing a phase."""

    def __init__(self, kind: int, phase: Phase, y0: np.ndarray, cond: np.ndarray, constantspecies: Sequence[str] = ()):
        self.kind, self.phase = kind, phase
        self.y0 = np.ascontiguousarray(np.atleast_2d(y0), dtype=float)
        self.cond = np.ascontiguousarray(np.atleast_2d(cond), dtype=float)
        self.n = self.y0.shape[0]
        Ns, Nr = phase.Ns, phase.Nr
        self.Ny = self.y0.shape[1]
        self.constantspeciesinds = np.array([phase.spcdict[k] for k in constantspecies], dtype=np.int32)
        self.parameterindexes = [1, Ns + Nr]
        self.efficiencyinds = list(phase.efficiencyinds)
        if kind == L.DOMAIN_CONSTANT_V:
            self.indexes = [1, Ns, Ns + 1, Ns + 2]
            self.thermovariabledict = {"T": Ns + 1, "P": Ns + 2}
        elif kind == L.DOMAIN_CONSTANT_P:
            self.indexes = [1, Ns, Ns + 1, Ns + 2]
            self.thermovariabledict = {"T": Ns + 1, "V": Ns + 2}
        elif kind == L.DOMAIN_CONSTANT_TP:
            self.indexes = [1, Ns, Ns + 1]
            self.thermovariabledict = {"V": Ns + 1}
        else:
            self.indexes = [1, Ns]
            self.thermovariabledict = {}
        self.p: Optional[np.ndarray] = None
        self._p_default: Optional[np.ndarray] = None

    # -- names mirrored from the reference structs ---------------------------------------------
    @property
    def T(self):
        return None if self.kind in (L.DOMAIN_CONSTANT_V, L.DOMAIN_CONSTANT_P) else _scalar(self.cond[:, 0])

    @property
    def P(self):
        if self.kind == L.DOMAIN_CONSTANT_P:
            return _scalar(self.cond[:, 0])
        return _scalar(self.cond[:, 1]) if self.kind == L.DOMAIN_CONSTANT_TP else None

    @property
    def V(self):
        if self.kind == L.DOMAIN_CONSTANT_V:
            return _scalar(self.cond[:, 0])
        return _scalar(self.cond[:, 1]) if self.kind == L.DOMAIN_CONSTANT_TV else None

    @property
    def A(self):
        return _scalar(self.cond[:, 1]) if self.kind == L.DOMAIN_CONSTANT_TAPHI else None

    def _p_for_device(self, p):
        """None when p is the constructor default (the device then evaluates its own tables, the
        'nothermochg && nokfchg' branch of calcthermo, src/Domain.jl:941-946,1532-1533)."""
        if p is None:
            return None, 0
        p = np.ascontiguousarray(p, dtype=float)
        if self._p_default is not None and p.shape == self._p_default.shape and np.array_equal(p, self._p_default):
            return None, 0
        if self.kind == L.DOMAIN_COUPLED_TP_TAPHI:
            raise NotImplementedError("non-default parameters for coupled gas-surface systems (DESIGN.md, 'next')")
        Np = self.phase.Ns + self.phase.Nr
        if p.ndim == 1:
            assert p.size == Np
            return p, 0
        assert p.shape == (self.n, Np)
        return p, Np

    # -- batched hot-path calls ------------------------------------------------------------------
    def _cond_for(self, n: int) -> np.ndarray:
        """Per-reactor constants for n states: the domain's own rows, or its single row repeated (one condition, many states)."""
        if self.cond.shape[0] == n:
            return self.cond
        if self.cond.shape[0] == 1:
            return np.ascontiguousarray(np.repeat(self.cond, n, axis=0))
        raise ValueError(f"{n} states for a domain of {self.cond.shape[0]} reactors")

    def _interfaces_for(self, interfaces, n: int):
        """(array, stride) of the boundary interfaces for n states: a packed row / matrix (rmsb200.interfaces.pack_interfaces) or a
        list of interface objects attached to this domain."""
        if interfaces is None or (isinstance(interfaces, (list, tuple)) and len(interfaces) == 0):
            return None, 0
        if not isinstance(interfaces, np.ndarray):
            from .interfaces import pack_interfaces
            interfaces = pack_interfaces(self, interfaces)
        a = np.ascontiguousarray(interfaces, dtype=float)
        if a.ndim == 1:
            return a, 0
        if a.shape[0] != n:
            raise ValueError("one interface row per reactor")
        return a, a.shape[1]

    def _extras(self, interfaces, n: int, t=None):
        """(rmsb_extras or None, arrays to keep alive) for n states: boundary interfaces, the domain's time profile (Parametrized*
        domains: self.profile, one table or one per reactor) and the evaluation times."""
        fa, fs = self._interfaces_for(interfaces, n)
        prof = getattr(self, "profile", None)
        if fa is None and prof is None:
            return None, ()
        x = L.Extras()
        x.struct_size = C.sizeof(L.Extras)
        keep = [fa]
        if fa is not None:
            x.interfaces, x.if_stride = fa.ctypes.data, fs
        if prof is not None:
            pa = np.ascontiguousarray(prof, dtype=float)
            if pa.ndim == 2 and pa.shape[0] == 1:
                pa = np.ascontiguousarray(pa[0])
            if pa.ndim == 2 and pa.shape[0] != n:
                raise ValueError("one profile row per reactor")
            x.profile, x.profile_stride = pa.ctypes.data, (0 if pa.ndim == 1 else pa.shape[1])
            keep.append(pa)
            if t is not None:
                ta = np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=float), (n,)))
                x.t = ta.ctypes.data
                keep.append(ta)
        return x, keep

    def dydt(self, y=None, p=None, interfaces=None, t=None):
        """dydtreactor! for every reactor of the ensemble (src/Reactor.jl:653-660), with boundary interfaces if given; t: evaluation
        time(s), only looked at by the Parametrized* domains."""
        y = self.y0 if y is None else np.ascontiguousarray(np.atleast_2d(y), dtype=float)
        pp, ps = self._p_for_device(p)
        out = np.empty_like(y)
        cs = self.constantspeciesinds
        cond = self._cond_for(y.shape[0])
        x, keep = self._extras(interfaces, y.shape[0], t)
        if x is not None:
            L.check(L.lib().rmsb_dydt_batch_ex(self.phase.handle, self.kind, y.shape[0], L.ptr(y), L.ptr(cond), L.ptr(pp), ps,
                                               L.ptr(cs) if cs.size else None, cs.size, C.byref(x), L.ptr(out)))
            return out
        L.check(L.lib().rmsb_dydt_batch(self.phase.handle, self.kind, y.shape[0], L.ptr(y), L.ptr(cond), L.ptr(pp), ps,
                                        L.ptr(cs) if cs.size else None, cs.size, L.ptr(out)))
        return out

    def jacobiany(self, y=None, p=None, interfaces=None, t=None):
        """jacobiany! for every reactor -> [n][Ny][Ny] with J[r, eqn, var] (src/Domain.jl:2204-3091)."""
        y = self.y0 if y is None else np.ascontiguousarray(np.atleast_2d(y), dtype=float)
        pp, ps = self._p_for_device(p)
        out = np.empty((y.shape[0], self.Ny * self.Ny))
        cs = self.constantspeciesinds
        cond = self._cond_for(y.shape[0])
        x, keep = self._extras(interfaces, y.shape[0], t)
        if x is not None:
            L.check(L.lib().rmsb_jac_batch_ex(self.phase.handle, self.kind, y.shape[0], L.ptr(y), L.ptr(cond), L.ptr(pp), ps,
                                              L.ptr(cs) if cs.size else None, cs.size, C.byref(x), L.ptr(out)))
        else:
            L.check(L.lib().rmsb_jac_batch(self.phase.handle, self.kind, y.shape[0], L.ptr(y), L.ptr(cond), L.ptr(pp), ps,
                                           L.ptr(cs) if cs.size else None, cs.size, L.ptr(out)))
        return out.reshape(y.shape[0], self.Ny, self.Ny).transpose(0, 2, 1)   # column-major -> [eqn, var]

    def jacobianp(self, y=None, p=None, interfaces=None, t=None):
        """jacobianp! for every reactor -> [n][Ny][Ns+Nr] with Jp[r, eqn, par] (src/Domain.jl:3095-3485); `interfaces`: boundary
        interfaces as for dydt / jacobiany (only an Inlet on ConstantV / ConstantP adds terms, src/Domain.jl:3421-3430)."""
        y = self.y0 if y is None else np.ascontiguousarray(np.atleast_2d(y), dtype=float)
        pp, ps = self._p_for_sens(p)
        Np = self.phase.Ns + self.phase.Nr
        out = np.empty((y.shape[0], Np * self.Ny))
        cs = self.constantspeciesinds
        xtr, keep = self._extras(interfaces, y.shape[0], t)          # boundary interfaces and / or the profile of a Parametrized* domain
        if xtr is not None:
            if getattr(self, "profile", None) is not None:
                pp, ps = self._p_for_device(p)                        # variable-k parameters [dG; f] whatever the layout
            L.check(L.lib().rmsb_jacp_batch_ex(self.phase.handle, self.kind, y.shape[0], L.ptr(y), L.ptr(self._cond_for(y.shape[0])), L.ptr(pp), ps,
                                               L.ptr(cs) if cs.size else None, cs.size, C.byref(xtr), L.ptr(out)))
        else:
            L.check(L.lib().rmsb_jacp_batch(self.phase.handle, self.kind, y.shape[0], L.ptr(y), L.ptr(self._cond_for(y.shape[0])), L.ptr(pp), ps,
                                            L.ptr(cs) if cs.size else None, cs.size, L.ptr(out)))
        return out.reshape(y.shape[0], Np, self.Ny).transpose(0, 2, 1)      # column-major -> [eqn, par]

    def _p_for_sens(self, p):
        """Parameter Jacobians of ConstantTV need p itself on the device (chain rule through the diffusion limit,
        src/Domain.jl:3290-3317), also when it is the constructor default."""
        if self.kind != L.DOMAIN_CONSTANT_TV:
            return self._p_for_device(p)
        p = self.p if p is None else p
        if p is None:
            raise ValueError("ConstantTV parameter Jacobian needs the parameter vector p")
        p = np.ascontiguousarray(p, dtype=float)
        return (p, 0) if p.ndim == 1 else (p, self.phase.Ns + self.phase.Nr)

    def rates(self, y=None, p=None):
        """(kfs, krevs) as calcthermo returns them -> two [n][Nr] arrays."""
        y = self.y0 if y is None else np.ascontiguousarray(np.atleast_2d(y), dtype=float)
        pp, ps = self._p_for_device(p)
        kf = np.empty((y.shape[0], self.phase.Nr))
        kr = np.empty_like(kf)
        L.check(L.lib().rmsb_rates_batch(self.phase.handle, self.kind, y.shape[0], L.ptr(y), L.ptr(self._cond_for(y.shape[0])), L.ptr(pp), ps,
                                         L.ptr(kf), L.ptr(kr)))
        return kf, kr


def _scalar(a):
    return float(a[0]) if a.size == 1 else a


def _parse(phase: Phase, initialconds: Dict[str, float], special: Sequence[str]):
    vals = {k: 0.0 for k in special}
    ns = np.zeros(phase.Ns)
    for key, val in initialconds.items():
        if key in vals:
            vals[key] = float(val)
        else:
            assert key in phase.spcdict, f"{key} not found in species list: {phase.names}"
            ns[phase.spcdict[key] - 1] = float(val)
    return vals, ns


def _finish_constk(dom: Domain, efficiency_to_one: bool):
    """p = [Gs; kfs] of the constant-k domains (src/Domain.jl:88-93,549-551,700-702)."""
    ph = dom.phase
    Gs = ph.gibbs(dom.cond[:, 0])
    kf, kr = dom.rates()
    dom.kfs, dom.krevs, dom.Gs = kf, kr, Gs
    kfp = kf.copy()
    if ph.diffusionlimited:
        # liquid phases store the rate coefficients BEFORE diffusion limitation (kfsnondiff, src/Domain.jl:549-551)
        kfp = np.empty_like(kf)
        L.check(L.lib().rmsb_kfs_batch(ph.handle, dom.kind, dom.n, L.ptr(dom.y0), L.ptr(dom.cond), L.ptr(kfp)))
        dom.kfsnondiff = kfp.copy()
    if efficiency_to_one:
        for ind in ph.efficiencyinds:
            kfp[:, ind - 1] = 1.0
    p = np.concatenate([Gs, kfp], axis=1)
    dom.p = p[0] if dom.n == 1 else p
    dom._p_default = dom.p.copy()


# ---- reference-style constructors: one condition from a Dict ------------------------------------
def ConstantVDomain(*, phase: Phase, initialconds: Dict[str, float], constantspecies: Sequence[str] = (), **_):
    """src/Domain.jl:120-170 -> (domain, y0, p); y0 = [ns; T; P], p = [zeros(Ns); ones(Nr)]."""
    assert phase.kind == L.PHASE_GAS
    v, ns = _parse(phase, initialconds, ("T", "P", "V"))
    T, P, V = v["T"], v["P"], v["V"]
    assert V != 0.0 or (T != 0.0 and P != 0.0)
    N = ns.sum()
    if V == 0.0:
        V = N * R * T / P
    elif T == 0.0:
        T = P * V / (R * N)
    elif P == 0.0:
        P = N * R * T / V
    else:
        raise ValueError("ConstantVDomain overspecified with T,P and V")
    dom = Domain(L.DOMAIN_CONSTANT_V, phase, np.concatenate([ns, [T, P]]), np.array([V]), constantspecies)
    dom.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
    dom._p_default = dom.p.copy()
    return dom, dom.y0[0].copy(), dom.p


def ConstantPDomain(*, phase: Phase, initialconds: Dict[str, float], constantspecies: Sequence[str] = (), **_):
    """src/Domain.jl:188-239 -> (domain, y0, p); adiabatic at constant pressure, y0 = [ns; T; V], p = [zeros(Ns); ones(Nr)]."""
    assert phase.kind == L.PHASE_GAS
    v, ns = _parse(phase, initialconds, ("T", "P", "V"))
    T, P, V = v["T"], v["P"], v["V"]
    assert P != 0.0 or (T != 0.0 and V != 0.0)
    N = ns.sum()
    if P == 0.0:
        P = N * R * T / V
    elif T == 0.0:
        T = P * V / (R * N)
    elif V == 0.0:
        V = N * R * T / P
    else:
        raise ValueError("ConstantPDomain overspecified with T,P and V")
    dom = Domain(L.DOMAIN_CONSTANT_P, phase, np.concatenate([ns, [T, V]]), np.array([P]), constantspecies)
    dom.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
    dom._p_default = dom.p.copy()
    return dom, dom.y0[0].copy(), dom.p


def ParametrizedVDomain(*, phase: Phase, initialconds: Dict[str, float], constantspecies: Sequence[str] = (), **_):
    """ParametrizedVDomain(; phase, initialconds, constantspecies) -> (domain, y0, p)  src/Domain.jl:331-411: adiabatic reactor whose
    volume follows a profile -- initialconds carries "ts" and "V" as arrays (the reference also accepts a function; here the profile
    is always the spline table the device evaluates).  y = [ns; T; P]; on the device this is the ConstantV domain with a time
    profile (include/rmsb200.h, rmsb_extras)."""
    from .profile import profile_value, spline_table
    assert phase.kind == L.PHASE_GAS, "ParametrizedVDomain needs an IdealGas phase"
    ic = dict(initialconds)
    tab = spline_table(ic.pop("ts"), ic.pop("V"))
    vals, ns = _parse(phase, ic, ("T", "P"))
    T, P = vals["T"], vals["P"]
    N = float(ns.sum())
    V0 = profile_value(tab, 0.0)
    if P == 0.0:
        P = N * R * T / V0
    elif T == 0.0:
        T = P * V0 / (R * N)
    y0 = np.concatenate([ns, [T, P]])
    dom = Domain(L.DOMAIN_CONSTANT_V, phase, y0, np.array([V0]), constantspecies)
    dom.profile = tab
    dom.indexes = [1, phase.Ns, phase.Ns + 1, phase.Ns + 2]
    dom.thermovariabledict = {"T": phase.Ns + 1, "P": phase.Ns + 2}
    dom.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
    dom._p_default = dom.p.copy()
    return dom, y0.copy(), dom.p.copy()


def ParametrizedTPDomain(*, phase: Phase, initialconds: Dict[str, float], constantspecies: Sequence[str] = (), **_):
    """ParametrizedTPDomain(; phase, initialconds, constantspecies) -> (domain, y0, p)  src/Domain.jl:257-320: temperature and pressure
    follow profiles ("ts", "T", "P" arrays), y = [ns; V], p = [zeros(Ns); ones(Nr)].  On the device: the ConstantTP layout with a
    T table followed by a P table; rate coefficients are evaluated on every call."""
    from .profile import profile_value, spline_table
    assert phase.kind == L.PHASE_GAS, "ParametrizedTPDomain needs an IdealGas phase"
    ic = dict(initialconds)
    ts = ic.pop("ts")
    tabT, tabP = spline_table(ts, ic.pop("T")), spline_table(ts, ic.pop("P"))
    _, ns = _parse(phase, ic, ())
    T0, P0 = profile_value(tabT, 0.0), profile_value(tabP, 0.0)
    V = float(ns.sum()) * R * T0 / P0
    y0 = np.concatenate([ns, [V]])
    dom = Domain(L.DOMAIN_CONSTANT_TP, phase, y0, np.array([T0, P0]), constantspecies)
    dom.profile = np.concatenate([tabT, tabP])
    dom.indexes = [1, phase.Ns, phase.Ns + 1]
    dom.thermovariabledict = {"V": phase.Ns + 1}
    dom.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
    dom._p_default = dom.p.copy()
    return dom, y0.copy(), dom.p.copy()


def ParametrizedTConstantVDomain(*, phase: Phase, initialconds: Dict[str, float], constantspecies: Sequence[str] = (), **_):
    """ParametrizedTConstantVDomain(; phase, initialconds, constantspecies) -> (domain, y0, p)  src/Domain.jl:563-633: a liquid whose
    temperature follows a profile ("ts", "T" arrays) at constant volume.  y = ns, p = [zeros(Ns); ones(Nr)].  On the device: the
    ConstantTV layout with a T table; thermo, viscosity, diffusivities and the diffusion-limited rate coefficients are evaluated on
    every call (src/Domain.jl:1372-1434)."""
    from .profile import profile_value, spline_table
    assert phase.kind == L.PHASE_LIQUID, "ParametrizedTConstantVDomain needs an IdealDiluteSolution phase"
    ic = dict(initialconds)
    assert "V" in ic and "P" not in ic, "ParametrizedTConstantVDomain needs V and cannot specify P"
    tab = spline_table(ic.pop("ts"), ic.pop("T"))
    vals, ns = _parse(phase, ic, ("V", "Phi"))
    y0 = ns.copy()
    dom = Domain(L.DOMAIN_CONSTANT_TV, phase, y0, np.array([profile_value(tab, 0.0), vals["V"]]), constantspecies)
    dom.profile = tab
    dom.indexes = [1, phase.Ns]
    dom.thermovariabledict = {}
    dom.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
    dom._p_default = dom.p.copy()
    return dom, y0.copy(), dom.p.copy()


def ParametrizedPDomain(*, phase: Phase, initialconds: Dict[str, float], constantspecies: Sequence[str] = (), **_):
    """ParametrizedPDomain(; phase, initialconds, constantspecies) -> (domain, y0, p)  src/Domain.jl:413-498: adiabatic reactor whose
    pressure follows a profile ("ts", "P" arrays).  y = [ns; T; V]; the ConstantP domain with a time profile on the device."""
    from .profile import profile_value, spline_table
    assert phase.kind == L.PHASE_GAS, "ParametrizedPDomain needs an IdealGas phase"
    ic = dict(initialconds)
    tab = spline_table(ic.pop("ts"), ic.pop("P"))
    vals, ns = _parse(phase, ic, ("T", "V"))
    T, V = vals["T"], vals["V"]
    N = float(ns.sum())
    P0 = profile_value(tab, 0.0)
    if V == 0.0:
        V = N * R * T / P0
    elif T == 0.0:
        T = P0 * V / (R * N)
    y0 = np.concatenate([ns, [T, V]])
    dom = Domain(L.DOMAIN_CONSTANT_P, phase, y0, np.array([P0]), constantspecies)
    dom.profile = tab
    dom.indexes = [1, phase.Ns, phase.Ns + 1, phase.Ns + 2]
    dom.thermovariabledict = {"T": phase.Ns + 1, "V": phase.Ns + 2}
    dom.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
    dom._p_default = dom.p.copy()
    return dom, y0.copy(), dom.p.copy()


def ConstantTPDomain(*, phase: Phase, initialconds: Dict[str, float], constantspecies: Sequence[str] = (), **_):
    """src/Domain.jl:42-102 -> (domain, y0, p); y0 = [ns; V], p = [Gs; kfs (efficiency reactions -> 1.0)]."""
    v, ns = _parse(phase, initialconds, ("T", "P"))
    T, P = v["T"], v["P"]
    assert T != 0.0 and P != 0.0
    V = ns.sum() * R * T / P
    dom = Domain(L.DOMAIN_CONSTANT_TP, phase, np.concatenate([ns, [V]]), np.array([T, P]), constantspecies)
    _finish_constk(dom, True)
    return dom, dom.y0[0].copy(), dom.p


def ConstantTVDomain(*, phase: Phase, initialconds: Dict[str, float], constantspecies: Sequence[str] = (), **_):
    """src/Domain.jl:500-560 -> (domain, y0, p); y0 = ns."""
    assert "P" not in initialconds, "ConstantTVDomain cannot specify P"
    v, ns = _parse(phase, initialconds, ("T", "V", "Phi"))
    assert v["T"] != 0.0 and v["V"] != 0.0 and v["Phi"] == 0.0
    dom = Domain(L.DOMAIN_CONSTANT_TV, phase, ns, np.array([v["T"], v["V"]]), constantspecies)
    _finish_constk(dom, False)
    return dom, dom.y0[0].copy(), dom.p


def ConstantTAPhiDomain(*, phase: Phase, initialconds: Dict[str, float], constantspecies: Sequence[str] = (), **_):
    """src/Domain.jl:660-711 -> (domain, y0, p); y0 = ns, 'V' is the area A (phi = 0 only)."""
    v, ns = _parse(phase, initialconds, ("T", "A", "Phi"))
    assert v["T"] != 0.0 and v["A"] != 0.0 and v["Phi"] == 0.0
    dom = Domain(L.DOMAIN_CONSTANT_TAPHI, phase, ns, np.array([v["T"], v["A"]]), constantspecies)
    _finish_constk(dom, False)
    return dom, dom.y0[0].copy(), dom.p


# ---- ensemble constructors: arrays of conditions -------------------------------------------------
def ConstantVEnsemble(phase: Phase, T, P, ns, constantspecies: Sequence[str] = ()) -> Domain:
    """n ConstantV reactors: T[n], P[n] initial temperature/pressure, ns[n][Ns] initial moles;
    V = N*R*T/P per reactor as in src/Domain.jl:143-145."""
    T, P = np.asarray(T, dtype=float), np.asarray(P, dtype=float)
    ns = np.atleast_2d(np.asarray(ns, dtype=float))
    V = ns.sum(axis=1) * R * T / P
    dom = Domain(L.DOMAIN_CONSTANT_V, phase, np.concatenate([ns, T[:, None], P[:, None]], axis=1), V[:, None], constantspecies)
    dom.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
    dom._p_default = dom.p.copy()
    return dom


def ConstantPEnsemble(phase: Phase, T, P, ns, constantspecies: Sequence[str] = ()) -> Domain:
    """n adiabatic constant-pressure reactors: T[n] initial temperature, P[n] pressure, ns[n][Ns]; V0 = N*R*T/P."""
    T, P = np.asarray(T, dtype=float), np.asarray(P, dtype=float)
    ns = np.atleast_2d(np.asarray(ns, dtype=float))
    V = ns.sum(axis=1) * R * T / P
    dom = Domain(L.DOMAIN_CONSTANT_P, phase, np.concatenate([ns, T[:, None], V[:, None]], axis=1), P[:, None], constantspecies)
    dom.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
    dom._p_default = dom.p.copy()
    return dom


def ConstantTPEnsemble(phase: Phase, T, P, ns, constantspecies: Sequence[str] = ()) -> Domain:
    T, P = np.asarray(T, dtype=float), np.asarray(P, dtype=float)
    ns = np.atleast_2d(np.asarray(ns, dtype=float))
    V = ns.sum(axis=1) * R * T / P
    return Domain(L.DOMAIN_CONSTANT_TP, phase, np.concatenate([ns, V[:, None]], axis=1), np.stack([T, P], axis=1), constantspecies)


def ConstantTVEnsemble(phase: Phase, T, V, ns, constantspecies: Sequence[str] = ()) -> Domain:
    T, V = np.asarray(T, dtype=float), np.asarray(V, dtype=float)
    return Domain(L.DOMAIN_CONSTANT_TV, phase, np.atleast_2d(np.asarray(ns, dtype=float)), np.stack([T, V], axis=1), constantspecies)
