"""Phases: the reference's flatten-once step (IdealGas / IdealDiluteSolution / IdealSurface ctors,
src/Phase.jl:40-136) producing the SoA tables that `rmsb_mech_create` uploads to the GPU.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from .mechanism import ElementaryReaction, Solvent, Species

_VECTORISABLE = ("Arrhenius", "PdepArrhenius", "Chebyshev")   # *vec constructors of Ratevec.jl, Troe excluded (Phase.jl:263)


def getkineticstype(kin: dict):
    """Grouping key of splitreactionsbykinetics (src/Calculators/Rate.jl:212-227, src/Phase.jl:227-243)."""
    t = kin["type"]
    if t == "PdepArrhenius":
        ps = tuple(math.trunc(p * 1000.0) / 1000.0 for p in kin["Ps"])
        if all(a["type"] == "Arrhenius" for a in kin["arrs"]):
            return ("PdepArrhenius", ps)
        return ("AbstRate", "PdepArrhenius", ps)
    if t == "Chebyshev":
        return ("Chebyshev", kin["Tmin"], kin["Tmax"], kin["Pmin"], kin["Pmax"], kin["coefs"].shape)
    return t


def _arr(a):
    return [a["A"], a["n"], a["Ea"]]


def _plog_block(kin: dict) -> List[float]:
    out = [float(len(kin["Ps"]))] + list(kin["Ps"])
    for a in kin["arrs"]:
        if a["type"] == "Arrhenius":
            out += [1.0] + _arr(a)
        elif a["type"] == "MultiArrhenius":
            out += [float(len(a["arrs"]))] + [v for q in a["arrs"] for v in _arr(q)]
        else:
            raise NotImplementedError(f"PdepArrhenius entry of type {a['type']}")
    return out


def pack_kinetics(kin: dict, vectorised: bool):
    """(RMSB_KIN_* tag, parameter list) in the layout documented in include/rmsb200.h."""
    t = kin["type"]
    if t == "Arrhenius":
        return (L.KIN_ARRHENIUS_VEC if vectorised else L.KIN_ARRHENIUS), _arr(kin)
    if t == "MultiArrhenius":
        return L.KIN_MULTIARRHENIUS, [float(len(kin["arrs"]))] + [v for q in kin["arrs"] for v in _arr(q)]
    if t == "PdepArrhenius":
        if vectorised:
            return L.KIN_PLOG_VEC, [float(len(kin["Ps"]))] + list(kin["Ps"]) + [v for q in kin["arrs"] for v in _arr(q)]
        return L.KIN_PLOG, _plog_block(kin)
    if t == "MultiPdepArrhenius":
        out = [float(len(kin["parrs"]))]
        for pk in kin["parrs"]:
            out += _plog_block(pk)
        return L.KIN_MULTIPLOG, out
    if t == "ThirdBody":
        return L.KIN_THIRDBODY, _arr(kin["arr"])
    if t == "Lindemann":
        return L.KIN_LINDEMANN, _arr(kin["arrhigh"]) + _arr(kin["arrlow"])
    if t == "Troe":
        return L.KIN_TROE, _arr(kin["arrhigh"]) + _arr(kin["arrlow"]) + [kin["a"], kin["T3"], kin["T1"], kin["T2"]]
    if t == "Chebyshev":
        nT, nP = kin["coefs"].shape
        return L.KIN_CHEBYSHEV_VEC, [kin["Tmin"], kin["Tmax"], kin["Pmin"], kin["Pmax"], float(nT), float(nP)] + \
            [float(v) for v in kin["coefs"].reshape(-1)]
    raise NotImplementedError(f"kinetics type {t!r} is outside the accelerated hot path (see DESIGN.md)")


class Phase:
    """Common flattening; use IdealGas / IdealDiluteSolution / IdealSurface."""

    def __init__(self, kind: int, species: Sequence[Species], reactions: Sequence[ElementaryReaction], *, name: str = "",
                 solvent: Optional[Solvent] = None, sitedensity: float = 0.0, diffusionlimited: bool = False, device: int = 0):
        self.kind, self.name, self.solvent, self.sitedensity = kind, name, solvent, float(sitedensity)
        self.diffusionlimited = bool(diffusionlimited)
        self.species = list(species)
        self.names = [s.name for s in self.species]
        self.spcdict = {s.name: i + 1 for i, s in enumerate(self.species)}    # 1-based like the reference
        # getveckinetics: group by kinetics type, vectorisable groups first (src/Phase.jl:42,248-281)
        keys, groups = [], []
        for r in reactions:
            k = getkineticstype(r.kinetics)
            if k in keys:
                groups[keys.index(k)].append(r)
            else:
                keys.append(k)
                groups.append([r])
        vec, other, self.veckineticsinds = [], [], []
        for k, g in zip(keys, groups):
            typ = k[0] if isinstance(k, tuple) else k
            if typ in _VECTORISABLE:
                vec += g
                self.veckineticsinds.append(len(vec))
            else:
                other += g
        self.nvec = len(vec)
        ordered = vec + other
        self.reactions = [ElementaryReaction(i + 1, r.reactants, r.products, r.kinetics, r.reversible, r.forwardable,
                                             r.electronchange) for i, r in enumerate(ordered)]
        self.Ns, self.Nr = len(self.species), len(self.reactions)
        # getreactionindices: rxnarray (reference: 8 x Nr; here row = reaction), src/Phase.jl:320-348
        self.rxnarray = np.zeros((self.Nr, 8), dtype=np.int32)
        for j, r in enumerate(self.reactions):
            if len(r.reactants) > 4 or len(r.products) > 4:
                raise ValueError("at most 4 reactants / products per reaction")
            for k, s in enumerate(r.reactants):
                self.rxnarray[j, k] = self.spcdict[s.name]
            for k, s in enumerate(r.products):
                self.rxnarray[j, 4 + k] = self.spcdict[s.name]
        self.reversibility = np.array([r.reversible for r in self.reactions], dtype=np.uint8)
        self.forwardability = np.array([r.forwardable for r in self.reactions], dtype=np.uint8)
        self.efficiencyinds = [r.index for r in self.reactions
                               if r.kinetics["type"] in ("ThirdBody", "Lindemann", "Troe") and r.kinetics["efficiencies"]]
        self._device = device
        self._vectorised = [j < self.nvec for j in range(self.Nr)]
        self._species_phase = None      # two-phase systems only (rmsb200/coupled.py)
        self._rxn_domain = None
        self._handle = None
        self._more_handles = {}
        self._build_desc()

    # -- SoA descriptor -----------------------------------------------------------------------
    def _build_desc(self):
        Ns, Nr = self.Ns, self.Nr
        maxr = max(len(s.thermo.polys) for s in self.species)
        self._nr = np.array([len(s.thermo.polys) for s in self.species], dtype=np.int32)
        self._nc = np.full((Ns, maxr), 7, dtype=np.int32)
        self._tmax = np.zeros((Ns, maxr))
        self._coefs = np.zeros((Ns, maxr, 9))
        for i, s in enumerate(self.species):
            for r, p in enumerate(s.thermo.polys):
                self._nc[i, r] = len(p.coefs)
                self._tmax[i, r] = p.Tmax
                self._coefs[i, r, :len(p.coefs)] = p.coefs
        ktype, koff, kpar, effptr, effspc, effval = [], [0], [], [0], [], []
        for j, r in enumerate(self.reactions):
            tag, pars = pack_kinetics(r.kinetics, vectorised=self._vectorised[j])
            ktype.append(tag)
            kpar += [float(v) for v in pars]
            koff.append(len(kpar))
            eff = r.kinetics.get("efficiencies") or {}
            for s in sorted(eff):      # reference iterates a Dict (unspecified order); ascending index here
                effspc.append(s)
                effval.append(eff[s])
            effptr.append(len(effspc))
        self._ktype = np.array(ktype, dtype=np.int32)
        self._koff = np.array(koff, dtype=np.int32)
        self._kpar = np.array(kpar if kpar else [0.0])
        self._effptr = np.array(effptr, dtype=np.int32)
        self._effspc = np.array(effspc if effspc else [0], dtype=np.int32)
        self._effval = np.array(effval if effval else [0.0])
        self._radius = np.array([s.radius for s in self.species])
        self._diffr = np.array([s.diffusion_r if s.diffusion_r is not None else -1.0 for s in self.species])
        d = L.MechDesc()
        d.struct_size = C.sizeof(L.MechDesc)
        d.phase_kind, d.n_species, d.n_reactions = self.kind, Ns, Nr
        d.sitedensity, d.diffusion_limited, d.nasa_max_ranges = self.sitedensity, int(self.diffusionlimited), maxr
        pi, pd, pu = (lambda a: a.ctypes.data_as(L._pi)), (lambda a: a.ctypes.data_as(L._pd)), (lambda a: a.ctypes.data_as(L._pu8))
        d.nasa_nranges, d.nasa_ncoef, d.nasa_tmax, d.nasa_coefs = pi(self._nr), pi(self._nc), pd(self._tmax), pd(self._coefs)
        d.rxnarray, d.kin_type, d.kin_offset, d.kin_params = pi(self.rxnarray), pi(self._ktype), pi(self._koff), pd(self._kpar)
        d.eff_ptr, d.eff_species, d.eff_value = pi(self._effptr), pi(self._effspc), pd(self._effval)
        d.reversible, d.forwardable = pu(self.reversibility), pu(self.forwardability)
        d.radius, d.diff_radius = pd(self._radius), pd(self._diffr)
        d.viscosity_kind = L.VISC_NONE
        if self._species_phase is not None:
            d.species_phase, d.rxn_domain = pi(self._species_phase), pi(self._rxn_domain)
        if self.solvent is not None:
            mu = self.solvent.mu
            if mu["type"] == "RiedelViscosity":
                d.viscosity_kind = L.VISC_RIEDEL
                for k, key in enumerate("ABCDE"):
                    d.viscosity_params[k] = mu[key]
            elif mu["type"] == "ConstantViscosity":
                d.viscosity_kind = L.VISC_CONSTANT
                d.viscosity_params[0] = mu["mu"]
            else:
                raise NotImplementedError(mu["type"])
        self.desc = d

    @property
    def handle(self):
        """Device tables, created on first use (`rmsb_mech_create`)."""
        if self._handle is None:
            h = C.c_void_p()
            L.check(L.lib().rmsb_mech_create(C.byref(self.desc), self._device, C.byref(h)))
            self._handle = h
        return self._handle

    def handle_on(self, device: int):
        """Handle of this mechanism on another GPU (tables are replicated per device; rmsb_solve_batch_multi)."""
        if device == self._device:
            return self.handle
        if device not in self._more_handles:
            h = C.c_void_p()
            L.check(L.lib().rmsb_mech_create(C.byref(self.desc), int(device), C.byref(h)))
            self._more_handles[device] = h
        return self._more_handles[device]

    def spec_source(self, domain_kind: int) -> str:
        """CUDA source of the kernels the library generates for this mechanism (large dydt / Jacobian batches of small
        all-Arrhenius gas mechanisms, csrc/rmsb_spec.cu).  Needs no device; raises when the mechanism does not qualify."""
        need = C.c_size_t(0)
        L.check(L.lib().rmsb_spec_source_desc(C.byref(self.desc), int(domain_kind), None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value + 1)
        L.check(L.lib().rmsb_spec_source_desc(C.byref(self.desc), int(domain_kind), buf, need.value + 1, C.byref(need)))
        return buf.value.decode()

    def specialized(self, domain_kind: int) -> bool:
        """True when large dydt / Jacobian batches of this mechanism run on kernels compiled for it (compiles them on first use)."""
        return bool(L.lib().rmsb_mech_specialized(self.handle, int(domain_kind)))

    def __del__(self):
        try:
            if L._lib is not None:
                for h in [self._handle, *self._more_handles.values()]:
                    if h is not None:
                        L._lib.rmsb_mech_destroy(h)
        except Exception:
            pass

    def gibbs(self, T, scalar_path=True):
        """Species Gibbs energies at temperatures T -> [n][Ns] (calcgibbs, src/PhaseState.jl:8-10)."""
        T = np.ascontiguousarray(np.atleast_1d(T), dtype=float)
        G = np.empty((T.size, self.Ns))
        L.check(L.lib().rmsb_gibbs_batch(self.handle, T.size, L.ptr(T), int(scalar_path), L.ptr(G)))
        return G


def IdealGas(species, reactions, name="", diffusionlimited=False, device=0):
    """IdealGas(species, reactions; name, diffusionlimited=false)  src/Phase.jl:40"""
    return Phase(L.PHASE_GAS, species, reactions, name=name, diffusionlimited=diffusionlimited, device=device)


def IdealDiluteSolution(species, reactions, solvent, name="", diffusionlimited=True, device=0):
    """IdealDiluteSolution(species, reactions, solvent; name, diffusionlimited=true)  src/Phase.jl:79"""
    return Phase(L.PHASE_LIQUID, species, reactions, name=name, solvent=solvent, diffusionlimited=diffusionlimited, device=device)


def IdealSurface(species, reactions, sitedensity, name="", diffusionlimited=False, device=0):
    """IdealSurface(species, reactions, sitedensity; name)  src/Phase.jl:118"""
    assert not diffusionlimited, "diffusionlimited=true not supported for IdealSurface"
    return Phase(L.PHASE_SURFACE, species, reactions, name=name, sitedensity=sitedensity, device=device)
