"""NumPy restatement of the RMS hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function cites the reference file:line (relative to /root/reference/) it follows.
All arithmetic is float64; every routine is written generically over the number type so
that the same code evaluates with complex arguments (complex-step differentiation stands
in for the reference's ForwardDiff dual numbers, src/Reactor.jl:1469-1477).

Indices are 0-based here; ``rxnarray`` keeps the reference's layout (8 x Nr, rows 0-3
reactants, 4-7 products) with -1 for "none" (reference: 1-based with 0 = none,
src/Phase.jl:320-348).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import yaml

# src/Constants.jl:1-7 -- NOTE R is 8.314 exactly, not CODATA.
R = 8.314
Na = 6.022e23
kB = R / Na
F = 96485.33212


# --------------------------------------------------------------------------------------
# Data model (src/Species.jl, src/Reaction.jl, src/Calculators/Thermo.jl:10-14,95-98)
# --------------------------------------------------------------------------------------
@dataclass
class NASAPoly:
    coefs: np.ndarray
    Tmin: float
    Tmax: float


@dataclass
class Species:
    name: str
    index: int  # 0-based position in its phase
    polys: List[NASAPoly]
    smiles: str = ""
    radius: float = 0.0
    diff_r: Optional[float] = None  # StokesDiffusivity radius (None = EmptyDiffusivity -> Inf)
    molecularweight: float = 0.0
    atomnums: Optional[Dict[str, int]] = None
    bondnum: int = -1


@dataclass
class Reaction:
    index: int
    reactants: List[str]
    products: List[str]
    kinetics: dict
    reversible: bool = True
    forwardable: bool = True
    electronchange: int = 0
    reactantinds: List[int] = field(default_factory=list)
    productinds: List[int] = field(default_factory=list)


@dataclass
class Solvent:
    name: str
    mu: dict  # {"type": "RiedelViscosity", A,B,C,D,E} or {"type":"ConstantViscosity","mu":..}

    def viscosity(self, T):
        m = self.mu
        if m["type"] == "RiedelViscosity":  # src/Calculators/Viscosity.jl:24
            return np.exp(m["A"] + m["B"] / T + m["C"] * np.log(T) + m["D"] * T ** m["E"])
        if m["type"] == "ConstantViscosity":
            return m["mu"]
        raise ValueError(m["type"])


# --------------------------------------------------------------------------------------
# Molecular information the reference obtains from RDKit at parse time
# (src/Parse.jl:127-148,191-205).  rdkit is not available: a small SMILES atom/bond counter
# (explicit-H molecule, like Chem.AddHs) covers the acyclic / simple-ring organic subset used
# by the fixtures.  Species containing elements outside Parse.jl:41-42's elementdict (e.g. Pt)
# get no molecular info, exactly as the reference's try/catch leaves them.
# --------------------------------------------------------------------------------------
_MCGOWAN = {  # m^3/mol, src/Parse.jl:49-57
    "H": 8.71e-6, "He": 6.75e-6, "Li": 22.23e-6, "C": 16.35e-6, "N": 14.39e-6, "O": 12.43e-6,
    "F": 10.47e-6, "Ne": 8.51e-6, "Si": 26.83e-6, "P": 24.87e-6, "S": 22.91e-6, "Cl": 20.95e-6,
    "Ar": 18.99e-6, "Br": 26.21e-6, "B": 18.31e-6,
}
_ATOMIC_WEIGHT = {"H": 1.008, "C": 12.011, "N": 14.007, "O": 15.999, "Ar": 39.948, "He": 4.003,
                  "Ne": 20.180, "S": 32.067, "Cl": 35.453, "F": 18.998}
_VALENCE = {"C": 4, "N": 3, "O": 2, "H": 1, "S": 2, "F": 1, "Cl": 1, "Br": 1, "B": 3, "P": 3}
_ELEMENTDICT = {"H", "Li", "C", "O", "N", "Cl", "S", "Ar", "Ne", "He", "P", "F", "Br", "I", "B"}


def smiles_atoms_bonds(smiles: str):
    """Atom counts, bond count (explicit-H molecule) and molecular weight [kg/mol]."""
    atoms = []  # [element, explicit_h or None, bond order sum]
    nbonds = 0
    stack, prev, bond = [], None, 1.0
    rings: Dict[str, Tuple[int, float]] = {}
    i = 0
    order = {"-": 1.0, "=": 2.0, "#": 3.0, "~": 1.0, ":": 1.5}
    while i < len(smiles):
        ch = smiles[i]
        if ch == "[":
            j = smiles.index("]", i)
            m = re.match(r"\d*([A-Z][a-z]?)(H\d*)?([+-]\d*)?", smiles[i + 1:j])
            el = m.group(1)
            nh = 0
            if m.group(2):
                nh = int(m.group(2)[1:]) if len(m.group(2)) > 1 else 1
            atoms.append([el, nh, 0.0])
            cur = len(atoms) - 1
            i = j + 1
        elif ch.isalpha():
            el = smiles[i:i + 2] if smiles[i:i + 2] in ("Cl", "Br") else ch
            atoms.append([el, None, 0.0])
            cur = len(atoms) - 1
            i += len(el)
        elif ch in order:
            bond = order[ch]
            i += 1
            continue
        elif ch == "(":
            stack.append(prev)
            i += 1
            continue
        elif ch == ")":
            prev = stack.pop()
            i += 1
            continue
        elif ch == ".":
            prev = None
            i += 1
            continue
        elif ch.isdigit():
            if ch in rings:
                other, b = rings.pop(ch)
                b = max(b, bond)
                atoms[other][2] += b
                atoms[prev][2] += b
                nbonds += 1
            else:
                rings[ch] = (prev, bond)
            bond = 1.0
            i += 1
            continue
        else:
            raise ValueError(f"unsupported SMILES character {ch!r} in {smiles!r}")
        if prev is not None:
            atoms[prev][2] += bond
            atoms[cur][2] += bond
            nbonds += 1
        prev, bond = cur, 1.0
    counts: Dict[str, int] = {}
    for el, nh, bsum in atoms:
        if el not in _ELEMENTDICT:
            raise KeyError(el)
        if nh is None:
            nh = max(0, int(round(_VALENCE.get(el, 0) - bsum)))
        counts[el] = counts.get(el, 0) + 1
        if nh:
            counts["H"] = counts.get("H", 0) + nh
            nbonds += nh
    mw = sum(_ATOMIC_WEIGHT[e] * n for e, n in counts.items()) / 1000.0
    return counts, nbonds, mw


def getspeciesradius(atomdict: Dict[str, int], nbonds: int) -> float:
    """McGowan radius, src/Parse.jl:191-205."""
    Vtot = 0.0
    for atm, n in atomdict.items():
        Vtot += n * _MCGOWAN[atm]
    Vtot -= nbonds * 6.56e-6
    V = Vtot / Na
    return (0.75 / math.pi * V) ** (1.0 / 3.0)


# --------------------------------------------------------------------------------------
# Parser: src/Parse.jl:281-429 (readinputyml), :208-245 (fcndict2obj)
# --------------------------------------------------------------------------------------
def _f(x):
    return float(x)


def _parse_kinetics(k: dict, spcindex: Dict[str, int]) -> dict:
    t = k["type"]
    out = {"type": t}
    if t in ("Arrhenius", "StickingCoefficient", "Arrheniusq"):
        out.update(A=_f(k["A"]), n=_f(k["n"]), Ea=_f(k["Ea"]))
        if t == "Arrheniusq":
            out["q"] = _f(k.get("q", 0.0))
    elif t == "MultiArrhenius":
        out["arrs"] = [_parse_kinetics(a, spcindex) for a in k["arrs"]]
    elif t == "PdepArrhenius":
        Ps = [_f(p) for p in k["Ps"]]
        arrs = [_parse_kinetics(a, spcindex) for a in k["arrs"]]
        # src/Calculators/Rate.jl:46 sorts Ps only (the fixtures are already sorted)
        out.update(Ps=sorted(Ps), arrs=arrs)
    elif t == "MultiPdepArrhenius":
        out["parrs"] = [_parse_kinetics(a, spcindex) for a in k["parrs"]]
    elif t in ("ThirdBody", "Lindemann", "Troe"):
        if t == "ThirdBody":
            out["arr"] = _parse_kinetics(k["arr"], spcindex)
        else:
            out["arrhigh"] = _parse_kinetics(k["arrhigh"], spcindex)
            out["arrlow"] = _parse_kinetics(k["arrlow"], spcindex)
        if t == "Troe":
            out.update(a=_f(k["a"]), T3=_f(k["T3"]), T1=_f(k["T1"]), T2=_f(k.get("T2", 0.0)))
        # src/Parse.jl:411-413: efficiencies stored by species index as (eff - 1)
        eff = k.get("efficiencies", {}) or {}
        out["efficiencies"] = {spcindex[name]: _f(v) - 1.0 for name, v in eff.items()}
    elif t == "Chebyshev":
        out.update(coefs=np.array(k["coefs"], dtype=float), Tmin=_f(k["Tmin"]), Tmax=_f(k["Tmax"]),
                   Pmin=_f(k["Pmin"]), Pmax=_f(k["Pmax"]))
    else:
        raise NotImplementedError(f"kinetics type {t}")
    return out


def readinput(fname: str) -> dict:
    """Restates readinputyml (src/Parse.jl:281-429): returns
    ``{phasename: {"Species": [...], "Reactions": [...]}, frozenset({p1,p2}): {"Reactions": [...]},
    "Solvents": [...]}``."""
    with open(fname) as fh:
        D = yaml.load(fh, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))
    out: dict = {}
    if "Solvents" in D and D["Solvents"]:
        out["Solvents"] = [Solvent(name=s["name"], mu={k: (_f(v) if k != "type" else v) for k, v in s["mu"].items()})
                           for s in D["Solvents"]]
    spcphase: Dict[str, Tuple[Species, str]] = {}
    for p in D["Phases"]:
        name = p["name"]
        spclist = []
        for idx, d in enumerate(p["Species"]):
            th = d["thermo"]
            if th["type"] != "NASA":
                raise NotImplementedError(f"thermo type {th['type']} (out of scope: SURVEY.md section 2 row 3)")
            polys = [NASAPoly(np.array([_f(c) for c in q["coefs"]]), _f(q["Tmin"]), _f(q["Tmax"])) for q in th["polys"]]
            sp = Species(name=d["name"], index=idx, polys=polys, smiles=d.get("smiles", ""))
            if "radius" in d:
                sp.radius = _f(d["radius"])
            if "molecularweight" in d:
                sp.molecularweight = _f(d["molecularweight"])
            # src/Parse.jl:306-346: molecular info from smiles when not given
            if sp.smiles and not ("atomnums" in d and "bondnum" in d and "molecularweight" in d):
                try:
                    sp.atomnums, sp.bondnum, sp.molecularweight = smiles_atoms_bonds(sp.smiles)
                except (KeyError, ValueError):
                    pass
            if "diffusion" in d:
                if d["diffusion"]["type"] == "StokesDiffusivity":
                    sp.diff_r = _f(d["diffusion"]["r"])
            elif sp.atomnums is not None and sp.bondnum >= 0:
                try:
                    r = getspeciesradius(sp.atomnums, sp.bondnum)
                    sp.diff_r = r
                    if "radius" not in d:
                        sp.radius = r
                except KeyError:
                    pass
            if sp.name in spcphase:
                raise ValueError("Two species in the file have the same name")
            spcphase[sp.name] = (sp, name)
            spclist.append(sp)
        out[name] = {"Species": spclist, "Reactions": []}
    for ridx, rx in enumerate(D["Reactions"]):
        phs = []
        for nm in list(rx["reactants"]) + list(rx["products"]):
            ph = spcphase[nm][1]
            if ph not in phs:
                phs.append(ph)
        spcindex = {nm: sp.index for nm, (sp, _) in spcphase.items()}
        r = Reaction(index=ridx, reactants=list(rx["reactants"]), products=list(rx["products"]),
                     kinetics=_parse_kinetics(rx["kinetics"], spcindex),
                     reversible=bool(rx.get("reversible", True)), forwardable=bool(rx.get("forwardable", True)),
                     electronchange=int(rx.get("electronchange", 0)))
        r.reactantinds = [spcphase[n][0].index for n in r.reactants]
        r.productinds = [spcphase[n][0].index for n in r.products]
        if len(phs) == 1:
            out[phs[0]]["Reactions"].append(r)
        else:
            out.setdefault(frozenset(phs), {"Reactions": []})["Reactions"].append(r)
    return out


# --------------------------------------------------------------------------------------
# Thermo: src/Calculators/Thermo.jl, src/Calculators/Thermovec.jl, src/Tools.jl:13-22
# --------------------------------------------------------------------------------------
def evalpoly(x, coefs):
    """Horner evaluation, src/Tools.jl:13-22."""
    out = 0.0
    for i in range(len(coefs) - 1, -1, -1):
        out = out + coefs[i]
        if i != 0:
            out = out * x
    return out


def selectPoly(polys: Sequence[NASAPoly], T) -> NASAPoly:
    """src/Calculators/Thermo.jl:101-111: first range with T <= Tmax, else the last."""
    Tr = T.real if isinstance(T, complex) else T
    for p in polys:
        if Tr <= p.Tmax:
            return p
    return polys[-1]


def getEnthalpy(poly: NASAPoly, T):
    c = poly.coefs
    if len(c) == 9:  # src/Calculators/Thermo.jl:87
        return ((-c[0] / T + c[1] * np.log(T)) / T + evalpoly(T, c[2:7] / np.arange(1, 6))) * R * T + c[7] * R
    return evalpoly(T, c[0:5] / np.arange(1, 6)) * R * T + c[5] * R  # :89


def getEntropy(poly: NASAPoly, T):
    c = poly.coefs
    if len(c) == 9:  # src/Calculators/Thermo.jl:77
        return ((-c[0] / (2 * T) - c[1]) / T + c[2] * np.log(T) + T * evalpoly(T, c[3:7] / np.arange(1, 5)) + c[8]) * R
    return (c[0] * np.log(T) + T * evalpoly(T, c[1:5] / np.arange(1, 5)) + c[6]) * R  # :79


def getHeatCapacity(poly: NASAPoly, T):
    c = poly.coefs
    if len(c) == 9:
        return evalpoly(T, c[0:7]) / T ** 2 * R
    return evalpoly(T, c[0:5]) * R


def getGibbs(poly: NASAPoly, T):
    """src/Calculators/Thermo.jl:6-8 (scalar thermo path, exact 1/k factors)."""
    return getEnthalpy(poly, T) - T * getEntropy(poly, T)


def calcHSCpdless_poly(c: np.ndarray, T):
    """src/Calculators/Thermo.jl:16-63 == Thermovec.jl:76-119 for one species.
    NOTE the literal 0.33333333333 (not 1/3) and the term ordering."""
    Tp0 = T
    Tp1 = T * T
    Tp2 = Tp1 * T
    Tp3 = Tp2 * T
    Tp4 = 1.0 / T
    Tp6 = np.log(T)
    if len(c) != 7:
        Tp5 = Tp4 * Tp4
        ct0 = c[0] * Tp5
        ct1 = c[1] * Tp4
        ct2 = c[2]
        ct3 = c[3] * Tp0
        ct4 = c[4] * Tp1
        ct5 = c[5] * Tp2
        ct6 = c[6] * Tp3
        cp = ct0 + ct1 + ct2 + ct3 + ct4 + ct5 + ct6
        h = -ct0 + Tp6 * ct1 + ct2 + 0.5 * ct3 + 0.33333333333 * ct4 + 0.25 * ct5 + 0.2 * ct6 + c[7] * Tp4
        s = -0.5 * ct0 - ct1 + Tp6 * ct2 + ct3 + 0.5 * ct4 + 0.33333333333 * ct5 + 0.25 * ct6 + c[8]
    else:
        ct2 = c[0]
        ct3 = c[1] * Tp0
        ct4 = c[2] * Tp1
        ct5 = c[3] * Tp2
        ct6 = c[4] * Tp3
        cp = ct2 + ct3 + ct4 + ct5 + ct6
        h = ct2 + 0.5 * ct3 + 0.33333333333 * ct4 + 0.25 * ct5 + 0.2 * ct6 + c[5] * Tp4
        s = Tp6 * ct2 + ct3 + 0.5 * ct4 + 0.33333333333 * ct5 + 0.25 * ct6 + c[6]
    return cp, h, s


# --------------------------------------------------------------------------------------
# Rate calculators: src/Calculators/Rate.jl, src/Calculators/Ratevec.jl
# --------------------------------------------------------------------------------------
def arrhenius(k, T):
    """Scalar form, src/Calculators/Rate.jl:16."""
    return k["A"] * T ** k["n"] * np.exp(-k["Ea"] / (R * T))


def arrhenius_vecform(A, n, Ea, T):
    """Vectorised form, src/Calculators/Ratevec.jl:24."""
    return A * np.exp(n * np.log(T) - Ea * (1.0 / (R * T)))


def getBoundingIndsSorted(el, x):
    """src/Tools.jl:26-40 (0-based; second index -1 = no interpolation)."""
    el = el.real if isinstance(el, complex) else el
    if el <= x[0]:
        return (0, -1)
    for i in range(len(x)):
        if x[i] >= el:
            if x[i] == el:
                return (i, -1)
            return (i - 1, i)
    return (len(x) - 1, -1)


def _log10(x):
    return np.log10(x)


def evalkinetics(k: dict, T, P=0.0, C=0.0, vecform=False):
    """Scalar rate calculators, src/Calculators/Rate.jl:10-210.  ``vecform`` evaluates Arrhenius
    leaves with the vectorised expression (Ratevec.jl:24) as PdepArrheniusvec does (:204-215)."""
    t = k["type"]
    if t == "Arrhenius":
        return arrhenius_vecform(k["A"], k["n"], k["Ea"], T) if vecform else arrhenius(k, T)
    if t == "StickingCoefficient":  # Rate.jl:26
        v = arrhenius(k, T)
        return v if (v.real if isinstance(v, complex) else v) < 1.0 else 1.0
    if t == "MultiArrhenius":  # Rate.jl:68-74
        out = 0.0
        for a in k["arrs"]:
            out = out + arrhenius(a, T)
        return out
    if t == "PdepArrhenius":  # Rate.jl:48-60
        i0, i1 = getBoundingIndsSorted(P, k["Ps"])
        if i1 == -1:
            return evalkinetics(k["arrs"][i0], T, vecform=vecform)
        highk = evalkinetics(k["arrs"][i1], T, vecform=vecform)
        lowk = evalkinetics(k["arrs"][i0], T, vecform=vecform)
        Plow, Phigh = k["Ps"][i0], k["Ps"][i1]
        return lowk * 10.0 ** (_log10(P / Plow) / _log10(Phigh / Plow) * _log10(highk / lowk))
    if t == "MultiPdepArrhenius":  # Rate.jl:82-88
        out = 0.0
        for pd in k["parrs"]:
            out = out + evalkinetics(pd, T, P)
        return out
    if t == "ThirdBody":  # Rate.jl:98
        return C * arrhenius(k["arr"], T)
    if t == "Lindemann":  # Rate.jl:109-114
        k0 = arrhenius(k["arrlow"], T)
        kinf = arrhenius(k["arrhigh"], T)
        Pr = k0 * C / kinf
        return kinf * Pr / (1.0 + Pr)
    if t == "Troe":  # Rate.jl:129-147
        k0 = arrhenius(k["arrlow"], T)
        kinf = arrhenius(k["arrhigh"], T)
        Pr = k0 * C / kinf
        if k["T1"] == 0.0 and k["T3"] == 0.0:
            Fv = 1.0
        else:
            Fcent = (1 - k["a"]) * np.exp(-T / k["T3"]) + k["a"] * np.exp(-T / k["T1"])
            if k["T2"] != 0.0:
                Fcent = Fcent + np.exp(-k["T2"] / T)
            d = 0.14
            n = 0.75 - 1.27 * _log10(Fcent)
            c = -0.4 - 0.67 * _log10(Fcent)
            Fv = 10.0 ** (_log10(Fcent) / (1 + ((_log10(Pr) + c) / (n - d * (_log10(Pr)))) ** 2))
        return kinf * (Pr / (1 + Pr)) * Fv
    if t == "Chebyshev":  # Rate.jl:150-209
        Tred = (2.0 / T - 1.0 / k["Tmin"] - 1.0 / k["Tmax"]) / (1.0 / k["Tmax"] - 1.0 / k["Tmin"])
        Pred = (2.0 * _log10(P) - _log10(k["Pmin"]) - _log10(k["Pmax"])) / (_log10(k["Pmax"]) - _log10(k["Pmin"]))
        kk = 0.0
        nT, nP = k["coefs"].shape
        for i in range(nT):
            for j in range(nP):
                kk = kk + k["coefs"][i, j] * _cheb(i, Tred) * _cheb(j, Pred)
        return 10.0 ** kk
    raise NotImplementedError(t)


def _cheb(n, x):
    """src/Calculators/Rate.jl:158-176."""
    if n == 0:
        return 1.0
    if n == 1:
        return x
    T0, T1 = 1.0, x
    Tn = 0.0
    for _ in range(n - 1):
        Tn = 2 * x * T1 - T0
        T0, T1 = T1, Tn
    return Tn


def chebyshev_vecform(k: dict, T, P):
    """src/Calculators/Ratevec.jl:89-100: loop order j (P) outer, i (T) inner, product of the
    two polynomial values formed first."""
    Tred = (2.0 / T - 1.0 / k["Tmin"] - 1.0 / k["Tmax"]) / (1.0 / k["Tmax"] - 1.0 / k["Tmin"])
    Pred = (2.0 * _log10(P) - _log10(k["Pmin"]) - _log10(k["Pmax"])) / (_log10(k["Pmax"]) - _log10(k["Pmin"]))
    kk = 0.0
    nT, nP = k["coefs"].shape
    for j in range(nP):
        for i in range(nT):
            kk = kk + k["coefs"][i, j] * (_cheb(i, Tred) * _cheb(j, Pred))
    return 10.0 ** kk


def calcdkdCeff(k: dict, T, Ceff):
    """src/Calculators/Rate.jl:232-262."""
    t = k["type"]
    if t == "ThirdBody":
        return arrhenius(k["arr"], T)
    k0 = arrhenius(k["arrlow"], T)
    kinf = arrhenius(k["arrhigh"], T)
    Pr = k0 * Ceff / kinf
    if t == "Lindemann" or (k["T1"] == 0.0 and k["T3"] == 0.0):
        return k0 / ((1.0 + Pr) * (1.0 + Pr))
    Fcent = (1 - k["a"]) * np.exp(-T / k["T3"]) + k["a"] * np.exp(-T / k["T1"])
    if k["T2"] != 0.0:
        Fcent = Fcent + np.exp(-k["T2"] / T)
    d = 0.14
    n = 0.75 - 1.27 * _log10(Fcent)
    c = -0.4 - 0.67 * _log10(Fcent)
    x = (_log10(Pr) + c) / (n - d * _log10(Pr))
    Fv = 10.0 ** (_log10(Fcent) / (1.0 + x * x))
    dFdCeff = (-np.log(10.0) * Fv * _log10(Fcent) / ((1.0 + x * x) * (1.0 + x * x)) * 2.0 * x
               * _log10(math.e) / Pr * k0 / kinf / (n - d * _log10(Pr)) * (1.0 + d * x))
    return k0 / ((1.0 + Pr) * (1.0 + Pr)) * Fv + kinf * Pr / (1.0 + Pr) * dFdCeff


def getkineticstype(kin: dict):
    """Grouping key, src/Calculators/Rate.jl:212-227."""
    t = kin["type"]
    if t == "PdepArrhenius":
        ps = tuple(math.trunc(p * 1000.0) / 1000.0 for p in kin["Ps"])
        if all(a["type"] == "Arrhenius" for a in kin["arrs"]):
            return ("PdepArrhenius", ps)
        return ("AbstRate", "PdepArrhenius", ps)
    if t == "Chebyshev":
        return ("Chebyshev", kin["Tmin"], kin["Tmax"], kin["Pmin"], kin["Pmax"], kin["coefs"].shape)
    return t


_VECTORISABLE = {"Arrhenius", "PdepArrhenius", "Chebyshev"}  # *vec ctors exported by Ratevec.jl, minus Troe (Phase.jl:263)


# --------------------------------------------------------------------------------------
# Phase flattening: src/Phase.jl:40-57 (IdealGas), :79-97, :118-136, :203-348
# --------------------------------------------------------------------------------------
class Phase:
    """IdealGas / IdealDiluteSolution / IdealSurface after the reference's flatten-once step."""

    def __init__(self, kind: str, species: List[Species], reactions: List[Reaction], *, name="",
                 solvent: Optional[Solvent] = None, sitedensity: float = 0.0, diffusionlimited: Optional[bool] = None):
        assert kind in ("gas", "liquid", "surface")
        self.kind, self.name, self.solvent, self.sitedensity = kind, name, solvent, sitedensity
        self.diffusionlimited = (kind == "liquid") if diffusionlimited is None else diffusionlimited
        self.species = species
        self.names = [s.name for s in species]
        self.spcdict = {s.name: i for i, s in enumerate(species)}
        # getveckinetics / splitreactionsbykinetics: src/Phase.jl:227-281
        keys, groups = [], []
        for r in reactions:
            key = getkineticstype(r.kinetics)
            if key in keys:
                groups[keys.index(key)].append(r)
            else:
                keys.append(key)
                groups.append([r])
        vec, other, self.vecgroups = [], [], []
        for key, grp in zip(keys, groups):
            typ = key[0] if isinstance(key, tuple) else key
            if typ in _VECTORISABLE:
                self.vecgroups.append((typ, len(vec), len(vec) + len(grp)))
                vec.extend(grp)
            else:
                other.extend(grp)
        self.nvec = len(vec)
        ordered = vec + other  # src/Phase.jl:42: vectorised first
        self.reactions = [Reaction(index=i, reactants=r.reactants, products=r.products, kinetics=r.kinetics,
                                   reversible=r.reversible, forwardable=r.forwardable, electronchange=r.electronchange,
                                   reactantinds=[self.spcdict[n] for n in r.reactants],
                                   productinds=[self.spcdict[n] for n in r.products])
                          for i, r in enumerate(ordered)]
        self.original_index = [r.index for r in ordered]
        Ns, Nr = len(species), len(self.reactions)
        self.Ns, self.Nr = Ns, Nr
        # getreactionindices: src/Phase.jl:320-348
        self.rxnarray = -np.ones((8, Nr), dtype=np.int64)
        for i, r in enumerate(self.reactions):
            for j, s in enumerate(r.reactantinds):
                self.rxnarray[j, i] = s
            for j, s in enumerate(r.productinds):
                self.rxnarray[j + 4, i] = s
        # getstoichmatrix: src/Phase.jl:203-222 (row index > 3 (1-based) counts as product)
        self.M = np.zeros((Nr, Ns))
        self.Nrp = np.zeros(Nr)
        for i in range(Nr):
            n = 0.0
            for j in range(8):
                ind = self.rxnarray[j, i]
                if ind < 0:
                    continue
                if j + 1 > 3:
                    self.M[i, ind] -= 1.0
                    n += 1.0
                else:
                    self.M[i, ind] += 1.0
                    n -= 1.0
            self.Nrp[i] = n
        self.reversibility = np.array([r.reversible for r in self.reactions], dtype=float)
        self.forwardability = np.array([r.forwardable for r in self.reactions], dtype=float)
        self.efficiencyinds = [r.index for r in self.reactions
                               if r.kinetics["type"] in ("ThirdBody", "Lindemann", "Troe") and len(r.kinetics["efficiencies"]) > 0]

    # -- thermo ------------------------------------------------------------------------
    def calcHSCpdless(self, T):
        """Vectorised thermo path: src/Calculators/Thermovec.jl:121-124,76-119 (value-equivalent to a
        per-species selectPoly)."""
        dt = complex if isinstance(T, complex) else float
        cp = np.zeros(self.Ns, dtype=dt)
        h = np.zeros(self.Ns, dtype=dt)
        s = np.zeros(self.Ns, dtype=dt)
        for i, sp in enumerate(self.species):
            cp[i], h[i], s[i] = calcHSCpdless_poly(selectPoly(sp.polys, T).coefs, T)
        return cp, h, s

    def calcgibbs(self, T):
        """Scalar thermo path: src/PhaseState.jl:8-10."""
        return np.array([getGibbs(selectPoly(sp.polys, T), T) for sp in self.species])

    def getC0(self, T):
        """src/Phase.jl:301-307."""
        return self.sitedensity if self.kind == "surface" else 1.0e5 / (R * T)

    # -- rate coefficients -------------------------------------------------------------
    def getkf(self, rxn: Reaction, T, P, C, ns, V):
        """src/PhaseState.jl:53-58 (efficiencies iterated in ascending species index; the reference
        iterates a Julia Dict, whose order is unspecified)."""
        kin = rxn.kinetics
        eff = kin.get("efficiencies")
        if eff:
            acc = 0.0
            for i in sorted(eff):
                acc = acc + ns[i] * eff[i]
            C = C + acc / V
        return evalkinetics(kin, T, P, C)

    def getkfs(self, T, P, C, ns, V):
        """src/PhaseState.jl:61-76."""
        dt = np.result_type(np.asarray(T).dtype, np.asarray(P).dtype, np.asarray(C).dtype, np.asarray(ns).dtype, float)
        kfs = np.zeros(self.Nr, dtype=dt)
        for typ, a, b in self.vecgroups:
            for i in range(a, b):
                kin = self.reactions[i].kinetics
                if typ == "Arrhenius":
                    kfs[i] = arrhenius_vecform(kin["A"], kin["n"], kin["Ea"], T)
                elif typ == "PdepArrhenius":
                    kfs[i] = evalkinetics(kin, T, P, vecform=True)
                else:
                    kfs[i] = chebyshev_vecform(kin, T, P)
        for i in range(self.nvec, self.Nr):
            kfs[i] = self.getkf(self.reactions[i], T, P, C, ns, V)
        return kfs

    def getKcs(self, T, Gs):
        """src/PhaseState.jl:165-167."""
        return np.exp(self.M @ (Gs / (R * T)) + self.Nrp * np.log(self.getC0(T)))

    def getKc(self, rxn: Reaction, T, Gs):
        """src/PhaseState.jl:106-129 (phi = 0)."""
        dG = 0.0
        r, p = rxn.reactantinds, rxn.productinds
        acc = Gs[r[0]]
        for i in r[1:]:
            acc = acc + Gs[i]
        dG = dG - acc
        acc = Gs[p[0]]
        for i in p[1:]:
            acc = acc + Gs[i]
        dG = dG + acc
        return np.exp(-(dG) / (R * T)) * self.getC0(T) ** (len(p) - len(r))

    def getDiffusiveRate(self, spcinds: Sequence[int], diffs):
        """src/PhaseState.jl:87-103 (1- and 2-body; the generalised >2-body form is a 'next' row)."""
        if len(spcinds) == 1:
            return np.inf
        if len(spcinds) == 2:
            a, b = spcinds
            return 4.0 * math.pi * (diffs[a] + diffs[b]) * (self.species[a].radius + self.species[b].radius) * Na
        # generalised Smoluchowski rate for N >= 3 bodies, src/PhaseState.jl:92-101
        D = np.array([diffs[i] for i in spcinds], dtype=float)
        n = len(spcinds)
        a = (3.0 * n - 5.0) / 2.0
        Dinv = 1.0 / D
        Dbar = 1.0 / np.cumsum(Dinv)[::-1]
        Dhat = D + Dbar
        deltaN = Dinv.sum() / sum(sum(1.0 / (D[i] * D[m]) for m in range(n - 1) if i > m) for i in range(1, n))
        rsum = sum(self.species[i].radius for i in spcinds)
        return np.prod(Dhat[1:] ** 1.5) * 4 * math.pi ** (a + 1) / math.gamma(a) * (rsum / math.sqrt(deltaN)) ** (2 * a) * Na ** (n - 1)

    def getkfkrev(self, rxn: Reaction, T, P, C, N, ns, Gs, diffs, V, kf=None, f=None):
        """src/PhaseState.jl:184-223."""
        if kf is None:
            kf = self.getkf(rxn, T, P, C, ns, V)
            if f is not None:
                kf = kf * f
        Kc = self.getKc(rxn, T, Gs)
        krev = kf / Kc
        if self.diffusionlimited:
            nr, npd = len(rxn.reactantinds), len(rxn.productinds)
            if nr == 1:
                if npd > 1:
                    krevdiff = self.getDiffusiveRate(rxn.productinds, diffs)
                    krev = krev * krevdiff / (krev + krevdiff)
                    kf = Kc * krev
            elif npd == 1:
                kfdiff = self.getDiffusiveRate(rxn.reactantinds, diffs)
                kf = kf * kfdiff / (kf + kfdiff)
                krev = kf / Kc
            elif npd == nr:
                kfdiff = self.getDiffusiveRate(rxn.reactantinds, diffs)
                krevdiff = self.getDiffusiveRate(rxn.productinds, diffs)
                kff = kf * kfdiff / (kf + kfdiff)
                krevr = krev * krevdiff / (krev + krevdiff)
                kfr = Kc * krevr
                if kff > kfr:
                    kf, krev = kfr, krevr
                else:
                    kf = kff
                    krev = kf / Kc
        kf = kf * float(rxn.forwardable)
        krev = krev * float(rxn.reversible)
        return kf, krev

    def getkfkrevs(self, T, P, C, N, ns, Gs, diffs, V, kfs=None):
        """src/PhaseState.jl:226-257."""
        if not self.diffusionlimited:
            if kfs is None:
                kfs = self.getkfs(T, P, C, ns, V)
            krev = kfs / self.getKcs(T, Gs)
        else:
            given = kfs
            dt = np.result_type(np.asarray(T).dtype, np.asarray(ns).dtype, np.asarray(Gs).dtype, float)
            kfs = np.zeros(self.Nr, dtype=dt)
            krev = np.zeros(self.Nr, dtype=dt)
            for i, rxn in enumerate(self.reactions):
                kfs[i], krev[i] = self.getkfkrev(rxn, T, P, C, N, ns, Gs, diffs, V,
                                                 kf=None if given is None else given[i])
        kfs = kfs * self.forwardability
        krev = krev * self.reversibility
        return kfs, krev

    def diffusivities(self, T, mu):
        """species.diffusion(T=T, mu=mu): src/Calculators/Diffusion.jl:8,19."""
        return np.array([np.inf if sp.diff_r is None else kB * T / (6 * math.pi * mu * sp.diff_r) for sp in self.species])


def IdealGas(species, reactions, name="", diffusionlimited=False):
    return Phase("gas", species, reactions, name=name, diffusionlimited=diffusionlimited)


def IdealDiluteSolution(species, reactions, solvent, name="", diffusionlimited=True):
    return Phase("liquid", species, reactions, name=name, solvent=solvent, diffusionlimited=diffusionlimited)


def IdealSurface(species, reactions, sitedensity, name="", diffusionlimited=False):
    assert not diffusionlimited
    return Phase("surface", species, reactions, name=name, sitedensity=sitedensity, diffusionlimited=False)


# --------------------------------------------------------------------------------------
# Mass-action source term: src/Reactor.jl:465-507 (and :556-598 with the extra factor)
# --------------------------------------------------------------------------------------
def addreactionratecontributions(dydt, rarray, cs, kfs, krevs, fac=None):
    for i in range(rarray.shape[1]):
        fR = kfs[i] * cs[rarray[0, i]]
        for j in (1, 2, 3):
            if rarray[j, i] < 0:
                break
            fR = fR * cs[rarray[j, i]]
        rR = krevs[i] * cs[rarray[4, i]]
        for j in (5, 6, 7):
            if rarray[j, i] < 0:
                break
            rR = rR * cs[rarray[j, i]]
        Rr = fR - rR
        if fac is not None:
            Rr = Rr * fac
        for j in range(4):
            if rarray[j, i] < 0:
                break
            dydt[rarray[j, i]] -= Rr
        for j in range(4, 8):
            if rarray[j, i] < 0:
                break
            dydt[rarray[j, i]] += Rr


def getrates(rarray, cs, kfs, krevs):
    """src/Reactor.jl:435-462."""
    out = np.zeros(rarray.shape[1], dtype=np.result_type(cs.dtype, kfs.dtype))
    for i in range(rarray.shape[1]):
        fR = kfs[i] * cs[rarray[0, i]]
        for j in (1, 2, 3):
            if rarray[j, i] < 0:
                break
            fR = fR * cs[rarray[j, i]]
        rR = krevs[i] * cs[rarray[4, i]]
        for j in (5, 6, 7):
            if rarray[j, i] < 0:
                break
            rR = rR * cs[rarray[j, i]]
        out[i] = fR - rR
    return out


# --------------------------------------------------------------------------------------
# Domains: src/Domain.jl
# --------------------------------------------------------------------------------------
@dataclass
class Thermo:
    """The 17-tuple returned by calcthermo (src/Domain.jl:916-...)."""
    ns: np.ndarray
    cs: np.ndarray
    T: float
    P: float
    V: float
    C: float
    N: float
    mu: float
    kfs: np.ndarray
    krevs: np.ndarray
    Hs: np.ndarray
    Us: np.ndarray
    Gs: np.ndarray
    diffs: np.ndarray
    Cvave: float
    cpdivR: np.ndarray
    phi: float = 0.0
    t: float = 0.0          # evaluation time (only the Parametrized* domains look at it)


_EMPTY = np.zeros(0)


def _initial(phase: Phase, initialconds: dict, special: Sequence[str]):
    vals = {k: 0.0 for k in special}
    ns = np.zeros(phase.Ns)
    for key, val in initialconds.items():
        if key in vals:
            vals[key] = float(val)
        else:
            assert key in phase.spcdict, f"{key} not found in species list"
            ns[phase.spcdict[key]] = float(val)
    return vals, ns


class _Domain:
    phase: Phase
    constantspeciesinds: List[int]

    @property
    def Ns(self):
        return self.phase.Ns

    def _constspc(self, constantspecies):
        return [self.phase.spcdict[k] for k in constantspecies]


class ConstantVDomain(_Domain):
    """src/Domain.jl:120-170 (ctor), :1064-1113 (calcthermo), :1901-1906 (derivatives). y = [ns; T; P]."""
    kind = "ConstantV"

    def __init__(self, phase: Phase, initialconds: dict, constantspecies=()):
        assert phase.kind == "gas"
        vals, ns = _initial(phase, initialconds, ("T", "P", "V"))
        T, P, V = vals["T"], vals["P"], vals["V"]
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
        self.phase, self.V = phase, V
        self.y0 = np.concatenate([ns, [T, P]])
        self.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
        self.constantspeciesinds = self._constspc(constantspecies)
        self.Ny = phase.Ns + 2
        self.cond = np.array([V])

    def calcthermo(self, y, t=0.0, p=None) -> Thermo:
        ph, Ns = self.phase, self.phase.Ns
        ns, T, P = y[:Ns], y[Ns], y[Ns + 1]
        V = self.V
        N = P * V / (R * T)
        cs = ns / V
        C = N / V
        cp, h, s = ph.calcHSCpdless(T)
        if p is not None:
            h = h + p[:Ns] / (R * T)
        Gs = (h - s) * (R * T)
        Us = (h - 1.0) * (R * T)
        Cv = np.dot(cp, ns)
        Cv = Cv * (R / N)
        Cv = Cv - R
        kfs, krevs = ph.getkfkrevs(T, P, C, N, ns, Gs, _EMPTY, V)
        if p is not None:
            kfs = kfs * p[Ns:Ns + ph.Nr]
            krevs = krevs * p[Ns:Ns + ph.Nr]
        return Thermo(ns, cs, T, P, V, C, N, 0.0, kfs, krevs, _EMPTY, Us, Gs, _EMPTY, Cv, cp)

    def derivatives(self, dydt, th: Thermo):
        Ns = self.phase.Ns
        dydt[Ns] = -np.dot(th.Us, dydt[:Ns]) / (th.N * th.Cvave)
        dydt[Ns + 1] = np.sum(dydt[:Ns]) * R * th.T / th.V + th.P / th.T * dydt[Ns]
        for i in self.constantspeciesinds:
            dydt[i] = 0.0


class ConstantPDomain(_Domain):
    """src/Domain.jl:173-239 (ctor), :1141-1192 (calcthermo), :1950-1953 (derivatives). y = [ns; T; V], adiabatic at constant P."""
    kind = "ConstantP"

    def __init__(self, phase: Phase, initialconds: dict, constantspecies=()):
        assert phase.kind == "gas"
        vals, ns = _initial(phase, initialconds, ("T", "P", "V"))
        T, P, V = vals["T"], vals["P"], vals["V"]
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
        self.phase, self.P = phase, P
        self.y0 = np.concatenate([ns, [T, V]])
        self.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
        self.constantspeciesinds = self._constspc(constantspecies)
        self.Ny = phase.Ns + 2
        self.cond = np.array([P])

    def calcthermo(self, y, t=0.0, p=None) -> Thermo:
        ph, Ns = self.phase, self.phase.Ns
        ns, T, V = y[:Ns], y[Ns], y[Ns + 1]
        P = self.P
        N = P * V / (R * T)
        cs = ns / V
        C = N / V
        cp, h, s = ph.calcHSCpdless(T)
        if p is not None:
            h = h + p[:Ns] / (R * T)
        Gs = (h - s) * (R * T)
        Hs = h * (R * T)
        Cv = np.dot(cp, ns)
        Cv = Cv * (R / N)
        Cv = Cv - R
        kfs, krevs = ph.getkfkrevs(T, P, C, N, ns, Gs, _EMPTY, V)
        if p is not None:
            kfs = kfs * p[Ns:Ns + ph.Nr]
            krevs = krevs * p[Ns:Ns + ph.Nr]
        return Thermo(ns, cs, T, P, V, C, N, 0.0, kfs, krevs, Hs, _EMPTY, Gs, _EMPTY, Cv, cp)

    def derivatives(self, dydt, th: Thermo):
        Ns = self.phase.Ns
        Cp = th.Cvave + R
        dydt[Ns] = -np.dot(th.Hs, dydt[:Ns]) / (th.N * Cp)
        dydt[Ns + 1] = np.sum(dydt[:Ns]) * R * th.T / th.P + dydt[Ns] * th.V / th.T
        for i in self.constantspeciesinds:
            dydt[i] = 0.0


class Spline:
    """getspline (src/Tools.jl:46-85): the reference fits a SmoothingSplines.jl smoothing spline with lambda = 1e-10 (third party, absent)
    and evaluates it with `_predict` (:54-85, restated in __call__).  With lambda -> 0 that fit is the interpolating NATURAL cubic
    spline, which is what is built here (gamma = second derivatives at the knots, zero at both ends); the difference to the
    reference's fit is O(1e-10) of the data -- parity unpinned at that level.  derivative(): analytic (the reference uses
    Calculus.derivative, a central finite difference, src/Domain.jl:2003,2047)."""

    def __init__(self, xs, vals):
        self.x = np.asarray(xs, dtype=float)
        self.g = np.asarray(vals, dtype=float)
        n = len(self.x)
        assert n >= 2 and np.all(np.diff(self.x) > 0)
        gam = np.zeros(n)
        if n > 2:
            h = np.diff(self.x)
            A = np.zeros((n - 2, n - 2))
            rhs = np.zeros(n - 2)
            for i in range(1, n - 1):
                A[i - 1, i - 1] = (h[i - 1] + h[i]) / 3.0
                if i > 1:
                    A[i - 1, i - 2] = h[i - 1] / 6.0
                if i < n - 2:
                    A[i - 1, i] = h[i] / 6.0
                rhs[i - 1] = (self.g[i + 1] - self.g[i]) / h[i] - (self.g[i] - self.g[i - 1]) / h[i - 1]
            gam[1:-1] = np.linalg.solve(A, rhs)
        self.gamma = gam

    def table(self):
        """[nt, ts, values, gamma]: the layout of the device profile (include/rmsb200.h, rmsb_extras.profile)."""
        return np.concatenate([[float(len(self.x))], self.x, self.g, self.gamma])

    def _piece(self, x):
        n = len(self.x)
        if x <= self.x[0]:
            h = self.x[1] - self.x[0]
            gp = (self.g[1] - self.g[0]) / h - h * self.gamma[1] / 6.0
            return self.g[0] - (self.x[0] - x) * gp, gp
        if x >= self.x[-1]:
            h = self.x[-1] - self.x[-2]
            gp = (self.g[-1] - self.g[-2]) / h + h * self.gamma[-2] / 6.0
            return self.g[-1] + (x - self.x[-1]) * gp, gp
        i = int(np.searchsorted(self.x, x, side="right")) - 1
        xl, xr, gl, gr, cl, cr = self.x[i], self.x[i + 1], self.g[i], self.g[i + 1], self.gamma[i], self.gamma[i + 1]
        h = xr - xl
        a, b = x - xl, xr - x
        w = (1 + a / h) * cr + (1 + b / h) * cl
        val = (a * gr + b * gl) / h - a * b * w / 6.0
        dv = (gr - gl) / h - ((b - a) * w + a * b * (cr - cl) / h) / 6.0
        return val, dv

    def __call__(self, x):
        return self._piece(float(x))[0]

    def derivative(self, x):
        return self._piece(float(x))[1]


class ParametrizedVDomain(ConstantVDomain):
    """src/Domain.jl:331-411 (ctor), :1219-1290 (calcthermo), :2002-2045 (derivatives): adiabatic reactor whose volume follows V(t).
    y = [ns; T; P]."""
    kind = "ParametrizedV"

    def __init__(self, phase: Phase, initialconds: dict, constantspecies=()):
        assert phase.kind == "gas"
        ic = dict(initialconds)
        ts, Vs = ic.pop("ts"), ic.pop("V")
        self.Vf = Spline(ts, Vs)
        vals, ns = _initial(phase, ic, ("T", "P"))
        T = vals["T"]
        N = ns.sum()
        V0 = self.Vf(0.0)
        P = vals["P"] if vals["P"] != 0.0 else N * R * T / V0
        if T == 0.0:
            T = P * V0 / (R * N)
        self.phase, self.V = phase, V0
        self.y0 = np.concatenate([ns, [T, P]])
        self.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
        self.constantspeciesinds = self._constspc(constantspecies)
        self.Ny = phase.Ns + 2
        self.cond = np.array([V0])

    def calcthermo(self, y, t=0.0, p=None) -> Thermo:
        self.V = self.Vf(t)
        th = ConstantVDomain.calcthermo(self, y, t, p)
        th.t = t
        return th

    def derivatives(self, dydt, th: Thermo):
        Ns = self.phase.Ns
        dVdt = self.Vf.derivative(th.t)
        dydt[Ns] = (-np.dot(th.Us, dydt[:Ns]) - th.P * dVdt) / (th.N * th.Cvave)
        dydt[Ns + 1] = np.sum(dydt[:Ns]) * R * th.T / th.V + dydt[Ns] * th.P / th.T - th.P / th.V * dVdt
        for i in self.constantspeciesinds:
            dydt[i] = 0.0


class ParametrizedPDomain(ConstantPDomain):
    """src/Domain.jl:413-498 (ctor), :1292-1368 (calcthermo), :2046-2095 (derivatives): adiabatic reactor whose pressure follows P(t).
    y = [ns; T; V]."""
    kind = "ParametrizedP"

    def __init__(self, phase: Phase, initialconds: dict, constantspecies=()):
        assert phase.kind == "gas"
        ic = dict(initialconds)
        ts, Ps = ic.pop("ts"), ic.pop("P")
        self.Pf = Spline(ts, Ps)
        vals, ns = _initial(phase, ic, ("T", "V"))
        T, V = vals["T"], vals["V"]
        N = ns.sum()
        P0 = self.Pf(0.0)
        if V == 0.0:
            V = N * R * T / P0
        elif T == 0.0:
            T = P0 * V / (R * N)
        self.phase, self.P = phase, P0
        self.y0 = np.concatenate([ns, [T, V]])
        self.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
        self.constantspeciesinds = self._constspc(constantspecies)
        self.Ny = phase.Ns + 2
        self.cond = np.array([P0])

    def calcthermo(self, y, t=0.0, p=None) -> Thermo:
        self.P = self.Pf(t)
        th = ConstantPDomain.calcthermo(self, y, t, p)
        th.t = t
        return th

    def derivatives(self, dydt, th: Thermo):
        Ns = self.phase.Ns
        Cp = th.Cvave + R
        dPdt = self.Pf.derivative(th.t)
        dydt[Ns] = (-np.dot(th.Hs, dydt[:Ns]) + th.V * dPdt) / (th.N * Cp)
        dydt[Ns + 1] = np.sum(dydt[:Ns]) * R * th.T / th.P + dydt[Ns] * th.V / th.T - dPdt * th.V / th.P
        for i in self.constantspeciesinds:
            dydt[i] = 0.0


class ParametrizedTPDomain(_Domain):
    """src/Domain.jl:241-320 (ctor), :1437-1512 (calcthermo), :2002-2004 (derivatives), :2635-2725 (jacobiany!): temperature and pressure
    follow time profiles, y = [ns; V].  A variable-k domain (vector thermo path, k's on every call); p = [dG; f]."""
    kind = "ParametrizedTP"

    def __init__(self, phase: Phase, initialconds: dict, constantspecies=()):
        assert phase.kind == "gas"
        ic = dict(initialconds)
        ts, Ts, Ps = ic.pop("ts"), ic.pop("T"), ic.pop("P")
        self.Tf, self.Pf = Spline(ts, Ts), Spline(ts, Ps)
        _, ns = _initial(phase, ic, ())
        N = ns.sum()
        V = N * R * self.Tf(0.0) / self.Pf(0.0)
        self.phase = phase
        self.y0 = np.concatenate([ns, [V]])
        self.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
        self.constantspeciesinds = self._constspc(constantspecies)
        self.Ny = phase.Ns + 1
        self.cond = np.array([self.Tf(0.0), self.Pf(0.0)])

    def calcthermo(self, y, t=0.0, p=None) -> Thermo:
        ph, Ns = self.phase, self.phase.Ns
        T, P = self.Tf(t), self.Pf(t)
        ns, V = y[:Ns], y[Ns]
        N = P * V / (R * T)
        cs = ns / V
        C = N / V
        cp, h, s = ph.calcHSCpdless(T)
        if p is not None:
            h = h + p[:Ns] / (R * T)
        Gs = (h - s) * (R * T)
        kfs, krevs = ph.getkfkrevs(T, P, C, N, ns, Gs, _EMPTY, V)
        if p is not None:
            kfs = kfs * p[Ns:Ns + ph.Nr]
            krevs = krevs * p[Ns:Ns + ph.Nr]
        return Thermo(ns, cs, T, P, V, C, N, 0.0, kfs, krevs, _EMPTY, _EMPTY, Gs, _EMPTY, 0.0, _EMPTY, 0.0, t)

    def derivatives(self, dydt, th: Thermo):
        Ns = self.phase.Ns
        dydt[Ns] = (np.sum(dydt[:Ns]) * R * th.T / th.P + self.Tf.derivative(th.t) * th.V / th.T
                    - self.Pf.derivative(th.t) * th.V / th.P)
        for i in self.constantspeciesinds:
            dydt[i] = 0.0


class ParametrizedTConstantVDomain(_Domain):
    """src/Domain.jl:563-633 (ctor), :1372-1434 (calcthermo), jacobiany! of the constant-T family (:3049-3055): a liquid whose
    temperature follows a time profile at constant volume.  y = ns; thermo through the vector path, viscosity, diffusivities and
    the diffusion-limited kf / krev (per-reaction path) on every call; p = [dG; f]."""
    kind = "ParametrizedTConstantV"

    def __init__(self, phase: Phase, initialconds: dict, constantspecies=()):
        assert phase.kind == "liquid"
        ic = dict(initialconds)
        assert "V" in ic and "P" not in ic, "ParametrizedTConstantVDomain needs V and cannot specify P"
        self.Tf = Spline(ic.pop("ts"), ic.pop("T"))
        vals, ns = _initial(phase, ic, ("V", "Phi"))
        self.phase, self.V, self.phi = phase, vals["V"], vals["Phi"]
        self.y0 = ns.copy()
        self.p = np.concatenate([np.zeros(phase.Ns), np.ones(phase.Nr)])
        self.constantspeciesinds = self._constspc(constantspecies)
        self.Ny = phase.Ns
        self.cond = np.array([self.Tf(0.0), self.V])

    def calcthermo(self, y, t=0.0, p=None) -> Thermo:
        ph, Ns = self.phase, self.phase.Ns
        T = self.Tf(t)
        ns = y[:Ns]
        N = np.sum(ns)
        cs = ns / self.V
        C = N / self.V
        P = 1.0e8
        mu = ph.solvent.viscosity(T)
        cp, h, s = ph.calcHSCpdless(T)
        if p is not None:
            h = h + p[:Ns] / (R * T)
        Gs = (h - s) * (R * T)
        diffs = ph.diffusivities(T, mu) if ph.diffusionlimited else _EMPTY
        kfs, krevs = ph.getkfkrevs(T, P, C, N, ns, Gs, diffs, self.V)
        if p is not None:
            kfs = kfs * p[Ns:Ns + ph.Nr]
            krevs = krevs * p[Ns:Ns + ph.Nr]
        return Thermo(ns, cs, T, P, self.V, C, N, mu, kfs, krevs, _EMPTY, _EMPTY, Gs, diffs, 0.0, _EMPTY, self.phi, t)

    def derivatives(self, dydt, th: Thermo):
        for i in self.constantspeciesinds:
            dydt[i] = 0.0


class _ConstantKDomain(_Domain):
    def _setup_k(self, T, P, C, N, ns, V):
        ph = self.phase
        self.Gs = ph.calcgibbs(T)
        if ph.solvent is not None:
            self.mu = ph.solvent.viscosity(T)
        else:
            self.mu = 0.0
        self.diffs = ph.diffusivities(T, self.mu) if ph.diffusionlimited else _EMPTY
        self.kfs, self.krevs = ph.getkfkrevs(T, P, C, N, ns, self.Gs, self.diffs, V)


class ConstantTPDomain(_ConstantKDomain):
    """src/Domain.jl:42-102 (ctor), :916-966 (calcthermo), :1863-1867 (derivatives). y = [ns; V]."""
    kind = "ConstantTP"

    def __init__(self, phase: Phase, initialconds: dict, constantspecies=()):
        vals, ns = _initial(phase, initialconds, ("T", "P"))
        T, P = vals["T"], vals["P"]
        assert T != 0.0 and P != 0.0
        self.phase, self.T, self.P = phase, T, P
        N = ns.sum()
        C = P / (R * T)
        V = N * R * T / P
        self.y0 = np.concatenate([ns, [V]])
        self._setup_k(T, P, C, N, ns, V)
        kfsp = self.kfs.copy()
        for i in phase.efficiencyinds:
            kfsp[i] = 1.0
        self.p = np.concatenate([self.Gs, kfsp])
        self.constantspeciesinds = self._constspc(constantspecies)
        self.Ny = phase.Ns + 1
        self.cond = np.array([T, P])

    def calcthermo(self, y, t=0.0, p=None) -> Thermo:
        ph, Ns = self.phase, self.phase.Ns
        ns, V = y[:Ns], y[Ns]
        N = self.P * V / (R * self.T)
        cs = ns / V
        C = N / V
        Gs = self.Gs
        if p is None or (np.array_equal(p[:Ns], self.Gs) and _kf_unchanged(self, p)):
            kfs, krevs = self.kfs.astype(np.result_type(self.kfs.dtype, y.dtype)), self.krevs.astype(np.result_type(self.krevs.dtype, y.dtype))
            kfps = None if p is None else p[Ns:Ns + ph.Nr]
        else:  # src/Domain.jl:948-961 -- thermo and/or k changed through p
            kfps = p[Ns:Ns + ph.Nr]
            Gs = p[:Ns]
            kfs = np.array(kfps, dtype=np.result_type(kfps.dtype, y.dtype))
            krevs = ph.getkfkrevs(self.T, self.P, C, N, ns, Gs, self.diffs, V, kfs=kfs.copy())[1]
        for ind in ph.efficiencyinds:  # src/Domain.jl:924,943
            kfs[ind], krevs[ind] = ph.getkfkrev(ph.reactions[ind], self.T, self.P, C, N, ns, Gs, self.diffs, V,
                                                f=None if kfps is None else kfps[ind])
        return Thermo(ns, cs, self.T, self.P, V, C, N, self.mu, kfs, krevs, _EMPTY, _EMPTY, _EMPTY, _EMPTY, 0.0, _EMPTY)

    def derivatives(self, dydt, th: Thermo):
        Ns = self.phase.Ns
        dydt[Ns] = np.sum(dydt[:Ns]) * R * th.T / th.P
        for i in self.constantspeciesinds:
            dydt[i] = 0.0


def _kf_unchanged(d, p):
    ph = d.phase
    kfps = p[ph.Ns:ph.Ns + ph.Nr]
    eff = ph.efficiencyinds
    return (np.count_nonzero(d.kfs != kfps) <= len(eff)) and all(kfps[i] == 1.0 for i in eff)


class ConstantTVDomain(_ConstantKDomain):
    """src/Domain.jl:500-560 (ctor), :1514-1561 (calcthermo), :1839-1842 (derivatives). y = ns."""
    kind = "ConstantTV"

    def __init__(self, phase: Phase, initialconds: dict, constantspecies=()):
        vals, ns = _initial(phase, initialconds, ("T", "V", "Phi"))
        T, V = vals["T"], vals["V"]
        assert "P" not in initialconds, "ConstantTVDomain cannot specify P"
        assert T != 0.0 and V != 0.0
        self.phase, self.T, self.V, self.phi = phase, T, V, vals["Phi"]
        N = ns.sum()
        P = 1.0e8
        C = N / V
        self.y0 = ns.copy()
        self._setup_k(T, P, C, N, ns, V)
        self.kfsnondiff = phase.getkfs(T, P, C, ns, V)
        self.p = np.concatenate([self.Gs, self.kfsnondiff])
        self.constantspeciesinds = self._constspc(constantspecies)
        self.Ny = phase.Ns
        self.cond = np.array([T, V])

    def calcthermo(self, y, t=0.0, p=None) -> Thermo:
        ph, Ns = self.phase, self.phase.Ns
        ns = y[:Ns]
        N = np.sum(ns)
        cs = ns / self.V
        C = N / self.V
        P = 1.0e8
        kfs, krevs = self.kfs, self.krevs
        if p is not None and not (np.array_equal(p[:Ns], self.Gs) and np.array_equal(p[Ns:Ns + ph.Nr], self.kfsnondiff)):
            kfs, krevs = ph.getkfkrevs(self.T, P, C, N, ns, p[:Ns], self.diffs, self.V, kfs=np.array(p[Ns:Ns + ph.Nr]))
        return Thermo(ns, cs, self.T, P, self.V, C, N, self.mu, kfs, krevs, _EMPTY, _EMPTY, _EMPTY, _EMPTY, 0.0, _EMPTY, self.phi)

    def derivatives(self, dydt, th: Thermo):
        for i in self.constantspeciesinds:
            dydt[i] = 0.0


class ConstantTAPhiDomain(_ConstantKDomain):
    """src/Domain.jl:660-711 (ctor), :1597-1633 (calcthermo). y = ns; 'V' is the area A."""
    kind = "ConstantTAPhi"

    def __init__(self, phase: Phase, initialconds: dict, constantspecies=()):
        vals, ns = _initial(phase, initialconds, ("T", "A", "Phi"))
        T, A = vals["T"], vals["A"]
        assert T != 0.0 and A != 0.0
        self.phase, self.T, self.A, self.phi = phase, T, A, vals["Phi"]
        self.V = A
        N = ns.sum()
        self.y0 = ns.copy()
        self._setup_k(T, 0.0, 0.0, N, ns, A)
        self.p = np.concatenate([self.Gs, self.kfs])
        self.constantspeciesinds = self._constspc(constantspecies)
        self.Ny = phase.Ns
        self.cond = np.array([T, A])

    def calcthermo(self, y, t=0.0, p=None) -> Thermo:
        ph, Ns = self.phase, self.phase.Ns
        ns = y[:Ns]
        N = np.sum(ns)
        cs = ns / self.A
        C = N / self.A
        kfs, krevs = self.kfs, self.krevs
        if p is not None:
            kfs = p[Ns:Ns + ph.Nr]
            if not np.array_equal(p[:Ns], self.Gs):
                krevs = ph.getkfkrevs(self.T, 0.0, C, N, ns, p[:Ns], self.diffs, self.A, kfs=np.array(kfs))[1]
        return Thermo(ns, cs, self.T, 0.0, self.A, C, N, self.mu, kfs, krevs, _EMPTY, _EMPTY, self.Gs, _EMPTY, 0.0, _EMPTY, self.phi)

    def derivatives(self, dydt, th: Thermo):
        for i in self.constantspeciesinds:
            dydt[i] = 0.0


# --------------------------------------------------------------------------------------
# RHS: src/Reactor.jl:653-660
# --------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------
# Boundary interfaces: src/Interface.jl:234-400 (structs), consumed by calcdomainderivatives!
# (src/Domain.jl:1839-1998) and by jacobiany! (:2220-2294, 2322-2448, 2493-2628, 3057-3088)
# --------------------------------------------------------------------------------------
def _flow(F, t):
    return F(t) if callable(F) else F


def _species_enthalpies(phase: Phase, T):
    return np.array([getEnthalpy(selectPoly(s.polys, T), T) for s in phase.species])


class Inlet:
    """Inlet(domain, conddict, F): molar flow F(t) of composition y (mole fractions) and enthalpy H (src/Interface.jl:234-263)."""

    def __init__(self, domain, conddict: dict, F):
        T, P = float(conddict.get("T", 0.0)), float(conddict.get("P", 0.0))
        y = np.zeros(domain.phase.Ns)
        for k, v in conddict.items():
            if k not in ("T", "P", "Hin"):
                y[domain.phase.spcdict[k]] = float(v)
        self.domain, self.y, self.F, self.T, self.P = domain, y / y.sum(), F, T, P
        if "Hin" in conddict:
            self.H = float(conddict["Hin"])
        else:
            assert T != 0.0 and P != 0.0, "T and P need to be provided if Hin is not provided for Inlet"
            self.H = float(np.dot(_species_enthalpies(domain.phase, T), self.y))


class Outlet:
    """Outlet(domain, F): molar flow F(t) of the reactor's own composition (src/Interface.jl:265-269)."""

    def __init__(self, domain, F):
        self.domain, self.F = domain, F


class VolumetricFlowRateInlet:
    """VolumetricFlowRateInlet(domain, conddict, Vin): volumetric flow of concentrations cs (src/Interface.jl:320-353)."""

    def __init__(self, domain, conddict: dict, Vin):
        T, P = float(conddict.get("T", 0.0)), float(conddict.get("P", 0.0))
        cs = np.zeros(domain.phase.Ns)
        for k, v in conddict.items():
            if k not in ("T", "P", "Hin"):
                cs[domain.phase.spcdict[k]] = float(v)
        self.domain, self.cs, self.Vin, self.T, self.P = domain, cs, Vin, T, P
        if "Hin" in conddict:
            self.Hpervolume = float(conddict["Hin"]) / cs.sum()
        else:
            assert T != 0.0 and P != 0.0
            self.Hpervolume = float(np.dot(_species_enthalpies(domain.phase, T), cs))


class VolumetricFlowRateOutlet:
    """src/Interface.jl:314-318"""

    def __init__(self, domain, Vout):
        self.domain, self.Vout = domain, Vout


class VolumeMaintainingOutlet:
    """Outlet whose flow keeps the volume constant (src/Interface.jl:356-365); evaluated after every other interface."""

    def __init__(self, domain):
        self.domain = domain


def apply_interfaces(domain, dydt, th: Thermo, t, interfaces):
    """The interface loops of calcdomainderivatives! (src/Domain.jl:1843-1860 generic, 1868-1898 ConstantTP, 1907-1947 ConstantV,
    1957-1997 ConstantP), after the constant-species zeroing."""
    Ns = domain.phase.Ns
    kind = {"ParametrizedV": "ConstantV", "ParametrizedP": "ConstantP", "ParametrizedTP": "ConstantTP"}.get(domain.kind, domain.kind)   # same interface terms (Domain.jl:2007-2094)
    T, P, V, N, ns = th.T, th.P, th.V, th.N, th.ns
    for it in interfaces:
        if it.domain is not domain:
            continue
        if isinstance(it, Inlet):
            flow = _flow(it.F, t)
            dydt[:Ns] += it.y * flow
            if kind == "ConstantTP":
                dydt[Ns] += flow * R * T / P
            elif kind == "ConstantV":
                dTdt = flow * (it.H - np.dot(th.Us, ns) / N) / (N * th.Cvave)
                dydt[Ns] += dTdt
                dydt[Ns + 1] += flow * R * T / V + P / T * dTdt
            elif kind == "ConstantP":
                dTdt = flow * (it.H - np.dot(th.Hs, ns) / N) / (N * (th.Cvave + R))
                dydt[Ns] += dTdt
                dydt[Ns + 1] += flow * R * T / P + dTdt * V / T
        elif isinstance(it, Outlet):
            flow = _flow(it.F, t)
            dydt[:Ns] -= flow * ns / N
            if kind == "ConstantTP":
                dydt[Ns] -= flow * R * T / P
            elif kind == "ConstantV":
                dTdt = (P * V / N * flow) / (N * th.Cvave)
                dydt[Ns] -= dTdt
                dydt[Ns + 1] -= flow * R * T / V + P / T * dTdt
            elif kind == "ConstantP":
                dydt[Ns + 1] -= flow * R * T / P
        elif isinstance(it, VolumetricFlowRateInlet):
            vin = _flow(it.Vin, t)
            dydt[:Ns] += vin * it.cs
            if kind == "ConstantTP":
                dydt[Ns] += vin
            elif kind == "ConstantV":
                dTdt = vin * (it.Hpervolume - np.dot(th.Us, ns) / V) / (N * th.Cvave)
                dydt[Ns] += dTdt
                dydt[Ns + 1] += vin * P / V + P / T * dTdt
            elif kind == "ConstantP":
                dTdt = vin * (it.Hpervolume - np.dot(th.Hs, ns) / V) / (N * (th.Cvave + R))
                dydt[Ns] += dTdt
                dydt[Ns + 1] += vin
        elif isinstance(it, VolumetricFlowRateOutlet):
            vout = _flow(it.Vout, t)
            dydt[:Ns] -= vout * ns / V
            if kind == "ConstantTP":
                dydt[Ns] -= vout
            elif kind == "ConstantV":
                dTdt = (P * vout) / (N * th.Cvave)
                dydt[Ns] -= dTdt
                dydt[Ns + 1] -= vout * P / V + P / T * dTdt
            elif kind == "ConstantP":
                dydt[Ns + 1] -= vout
    for it in interfaces:
        if isinstance(it, VolumeMaintainingOutlet) and it.domain is domain and kind in ("ConstantTP", "ConstantP"):
            iv = Ns if kind == "ConstantTP" else Ns + 1
            dVdt = dydt[iv]
            flow = P * dVdt / (R * T)
            dydt[:Ns] -= flow * ns / N
            dydt[iv] -= dVdt


def dydtreactor(domain, y, t=0.0, p=None, interfaces=()):
    y = np.asarray(y)
    th = domain.calcthermo(y, t, p)
    dt = np.result_type(y.dtype, th.kfs.dtype, np.asarray(th.V).dtype, float)
    dydt = np.zeros(len(y), dtype=dt)
    addreactionratecontributions(dydt, domain.phase.rxnarray, th.cs, th.kfs, th.krevs)
    dydt *= th.V
    domain.derivatives(dydt, th)
    if interfaces:
        apply_interfaces(domain, dydt, th, t, interfaces)
    return dydt


# --------------------------------------------------------------------------------------
# Analytic Jacobian: src/Domain.jl:2151-2454,3049-3055 ; src/Reactor.jl:874-1477
# --------------------------------------------------------------------------------------
def _spread(jac, deriv, slots, col, sign):
    for s in slots:
        jac[s, col] += sign * deriv


def _jacobianynswrtns(jac, rarray, i, cs, kf, krev):
    """Multiplicity-aware restatement of the hand-unrolled src/Reactor.jl:899-1366: for each distinct
    species s of one side with multiplicity m, deriv = m*k*prod(other concentrations); every slot of
    that side loses deriv, every slot of the other side gains it."""
    for side, other, k in ((range(0, 4), range(4, 8), kf), (range(4, 8), range(0, 4), krev)):
        slots = [rarray[j, i] for j in side if rarray[j, i] >= 0]
        oslots = [rarray[j, i] for j in other if rarray[j, i] >= 0]
        seen = []
        for s in slots:
            if s in seen:
                continue
            seen.append(s)
            m = slots.count(s)
            deriv = float(m) * k
            removed = False
            for q in slots:
                if q == s and not removed:
                    removed = True
                    continue
                deriv = deriv * cs[q]
            _spread(jac, deriv, slots, s, -1.0)
            _spread(jac, deriv, oslots, s, +1.0)


def _jacobianynswrtV(jac, Vind, rarray, i, cs, kf, krev):
    """src/Reactor.jl:1368-1413."""
    for side, other, k in ((range(0, 4), range(4, 8), kf), (range(4, 8), range(0, 4), krev)):
        slots = [rarray[j, i] for j in side if rarray[j, i] >= 0]
        oslots = [rarray[j, i] for j in other if rarray[j, i] >= 0]
        if len(slots) < 2:
            continue
        deriv = -(len(slots) - 1.0) * k
        for q in slots:
            deriv = deriv * cs[q]
        _spread(jac, deriv, slots, Vind, -1.0)
        _spread(jac, deriv, oslots, Vind, +1.0)


def _jacobianykderiv(jac, xind, dkfdx, dkrevdx, rarray, i, cs, V):
    """src/Reactor.jl:1418-1467."""
    for side, other, dk in ((range(0, 4), range(4, 8), dkfdx), (range(4, 8), range(0, 4), dkrevdx)):
        slots = [rarray[j, i] for j in side if rarray[j, i] >= 0]
        oslots = [rarray[j, i] for j in other if rarray[j, i] >= 0]
        deriv = dk
        for q in slots:
            deriv = deriv * cs[q]
        deriv = deriv * V
        _spread(jac, deriv, slots, xind, -1.0)
        _spread(jac, deriv, oslots, xind, +1.0)


def jacobianynsderiv(jac, domain, th: Thermo, Vind: Optional[int]):
    """src/Domain.jl:2169-2202 + :2151-2167."""
    ph = domain.phase
    for i in range(ph.Nr):
        _jacobianynswrtns(jac, ph.rxnarray, i, th.cs, th.kfs[i], th.krevs[i])
        if Vind is not None:
            _jacobianynswrtV(jac, Vind, ph.rxnarray, i, th.cs, th.kfs[i], th.krevs[i])
        if i in ph.efficiencyinds:
            kin = ph.reactions[i].kinetics
            Kc = th.kfs[i] / th.krevs[i]
            Ceff = th.C + sum(th.cs[s] * kin["efficiencies"][s] for s in sorted(kin["efficiencies"]))
            dkdCeff = calcdkdCeff(kin, th.T, Ceff)
            for s in sorted(kin["efficiencies"]):
                dkdni = dkdCeff * (kin["efficiencies"][s] / th.V)
                _jacobianykderiv(jac, s, dkdni, dkdni / Kc, ph.rxnarray, i, th.cs, th.V)
            if Vind is not None:
                dkdV = dkdCeff * (th.C - Ceff) / th.V
                _jacobianykderiv(jac, Vind, dkdV, dkdV / Kc, ph.rxnarray, i, th.cs, th.V)


def _jac_interfaces(jac, domain, th: Thermo, t, interfaces):
    """Analytic interface terms of the species columns (and of the volume column where the reference has them): src/Domain.jl:2220-2294
    (ConstantTP), 2322-2448 (ConstantV), 2493-2628 (ConstantP), 3057-3088 (ConstantTV / ConstantTAPhi).  VolumeMaintainingOutlet has
    no analytic terms in the reference."""
    Ns = domain.phase.Ns
    kind = {"ParametrizedV": "ConstantV", "ParametrizedP": "ConstantP", "ParametrizedTP": "ConstantTP"}.get(domain.kind, domain.kind)
    T, P, V, N, ns = th.T, th.P, th.V, th.N, th.ns
    for it in interfaces:
        if it.domain is not domain:
            continue
        if kind == "ConstantTP":
            if isinstance(it, Outlet):
                flow = _flow(it.F, t)
                for i in range(Ns):
                    jac[i, i] -= flow / N
                jac[:Ns, Ns] -= -flow * ns / (N * V)
            elif isinstance(it, VolumetricFlowRateOutlet):
                vout = _flow(it.Vout, t)
                for i in range(Ns):
                    jac[i, i] -= vout / V
                jac[:Ns, Ns] -= -vout * ns / (V * V)
        elif kind == "ConstantV":
            Cv = th.Cvave
            if isinstance(it, (Inlet, VolumetricFlowRateInlet)):
                if isinstance(it, Inlet):
                    flow = _flow(it.F, t)
                    dTdt = flow * (it.H - np.dot(th.Us, ns) / N) / (N * Cv)
                else:
                    flow = _flow(it.Vin, t)
                    dTdt = (flow * (it.Hpervolume - np.dot(th.Us, ns) / V)) / (N * Cv)
                for i in range(Ns):
                    dC = th.cpdivR[i] * R / N
                    dd = flow * (-th.Us[i] / N) / (N * Cv) - dTdt * (dC / Cv)
                    jac[Ns, i] += dd
                    jac[Ns + 1, i] += P / T * dd
            elif isinstance(it, (Outlet, VolumetricFlowRateOutlet)):
                if isinstance(it, Outlet):
                    flow = _flow(it.F, t)
                    dTdt = (P * V / N * flow) / (N * Cv)
                    diag = flow / N
                else:
                    flow = _flow(it.Vout, t)
                    dTdt = (P * flow) / (N * Cv)
                    diag = flow / V
                for i in range(Ns):
                    jac[i, i] -= diag
                    dC = th.cpdivR[i] * R / N
                    dd = -dTdt * (dC / Cv)
                    jac[Ns, i] -= dd
                    jac[Ns + 1, i] -= P / T * dd
        elif kind == "ConstantP":
            Cp = th.Cvave + R
            if isinstance(it, Inlet):
                flow = _flow(it.F, t)
                H = np.dot(th.Hs, ns) / N
                dTdt = flow * (it.H - H) / (N * Cp)
                for i in range(Ns):
                    dC = th.cpdivR[i] * R / N
                    dd = flow * (-th.Hs[i] / N) / (N * Cp) - dTdt * (dC / Cp)
                    jac[Ns, i] += dd
                    jac[Ns + 1, i] += V / T * dd
                ddV = flow * H / V / (N * Cp)
                jac[Ns, Ns + 1] += ddV
                jac[Ns + 1, Ns + 1] += V / T * ddV
            elif isinstance(it, Outlet):
                flow = _flow(it.F, t)
                for i in range(Ns):
                    jac[i, i] -= flow / N
                jac[:Ns, Ns + 1] -= -flow * ns / N / V
            elif isinstance(it, VolumetricFlowRateInlet):
                vin = _flow(it.Vin, t)
                dTdt = vin * (it.Hpervolume - np.dot(th.Hs, ns) / V) / (N * Cp)
                for i in range(Ns):
                    dC = th.cpdivR[i] * R / N
                    jac[Ns, i] += vin * (-th.Hs[i] / N) / (N * Cp) - dTdt * (dC / Cp)
            elif isinstance(it, VolumetricFlowRateOutlet):
                vout = _flow(it.Vout, t)
                for i in range(Ns):
                    jac[i, i] -= vout / V
                jac[:Ns, Ns + 1] -= -vout / (V * V) * ns
        else:   # ConstantTV / ConstantTAPhi (N = sum(ns))
            if isinstance(it, Outlet):
                flow = _flow(it.F, t)
                for i in range(Ns):
                    jac[i, i] -= flow / N
                    jac[i, :] += flow * ns[i] / (N * N)
            elif isinstance(it, VolumetricFlowRateOutlet):
                vout = _flow(it.Vout, t)
                for i in range(Ns):
                    jac[i, i] -= vout / V


def jacobiany(domain, y, t=0.0, p=None, cstep=1e-30, interfaces=()):
    """Analytic Jacobian jac[row = equation, col = variable], src/Domain.jl:2204-2454,3049-3055.
    The ConstantV T and P columns are the derivative of the complete RHS (reference: ForwardDiff,
    src/Reactor.jl:1469-1477; here: complex step with h = ``cstep``, exact to rounding)."""
    y = np.asarray(y, dtype=float)
    Ny = len(y)
    Ns = domain.phase.Ns
    jac = np.zeros((Ny, Ny))
    th = domain.calcthermo(y, t, p)
    if domain.kind in ("ConstantTP", "ParametrizedTP"):
        jacobianynsderiv(jac, domain, th, Ns)
        for i in range(Ns + 1):
            jac[Ns, i] = np.sum(jac[:Ns, i]) * R * th.T / th.P
        if domain.kind == "ParametrizedTP":                 # src/Domain.jl:2646
            jac[Ns, Ns] += domain.Tf.derivative(t) / th.T - domain.Pf.derivative(t) / th.P
        for ind in domain.constantspeciesinds:
            jac[ind, :] = 0.0
        _jac_interfaces(jac, domain, th, t, interfaces)
    elif domain.kind in ("ConstantV", "ParametrizedV"):
        jacobianynsderiv(jac, domain, th, None)
        dydt = np.zeros(Ny)
        addreactionratecontributions(dydt, domain.phase.rxnarray, th.cs, th.kfs, th.krevs)
        dydt *= th.V
        PdVdt = th.P * domain.Vf.derivative(t) if domain.kind == "ParametrizedV" else 0.0     # src/Domain.jl:2743-2746
        for i in range(Ns):
            dCvdni = th.cpdivR[i] * R / th.N
            jac[Ns, i] = (-np.dot(th.Us, jac[:Ns, i]) / (th.N * th.Cvave)
                          + (np.dot(th.Us, dydt[:Ns]) + PdVdt) / (th.N * th.Cvave * th.Cvave) * dCvdni)
            jac[Ns + 1, i] = np.sum(jac[:Ns, i]) * R * th.T / th.V + th.P / th.T * jac[Ns, i]
        for ind in domain.constantspeciesinds:
            jac[ind, :] = 0.0
        _jac_interfaces(jac, domain, th, t, interfaces)
        for col in (Ns, Ns + 1):
            yc = y.astype(complex)
            yc[col] += 1j * cstep
            jac[:, col] = dydtreactor(domain, yc, t, p, interfaces).imag / cstep
    elif domain.kind in ("ConstantP", "ParametrizedP"):      # src/Domain.jl:2456-2632 / 2874-2896: V column analytic, T column = derivative of the full RHS
        jacobianynsderiv(jac, domain, th, Ns + 1)
        dydt = np.zeros(Ny)
        addreactionratecontributions(dydt, domain.phase.rxnarray, th.cs, th.kfs, th.krevs)
        dydt *= th.V
        Cp = th.Cvave + R
        dPdt = domain.Pf.derivative(t) if domain.kind == "ParametrizedP" else 0.0
        dTdt = (-np.dot(th.Hs, dydt[:Ns]) + th.V * dPdt) / (th.N * Cp)
        for i in range(Ns):
            dCpdni = th.cpdivR[i] * R / th.N
            jac[Ns, i] = -np.dot(th.Hs, jac[:Ns, i]) / (th.N * Cp) - dTdt * (dCpdni / Cp)
            jac[Ns + 1, i] = np.sum(jac[:Ns, i]) * R * th.T / th.P + th.V / th.T * jac[Ns, i]
        jac[Ns, Ns + 1] = (-np.dot(th.Hs, jac[:Ns, Ns + 1]) + dPdt) / (th.N * Cp)
        jac[Ns + 1, Ns + 1] = np.sum(jac[:Ns, Ns + 1]) * R * th.T / th.P + dTdt / th.T + th.V / th.T * jac[Ns, Ns + 1] - dPdt / th.P
        for ind in domain.constantspeciesinds:
            jac[ind, :] = 0.0
        _jac_interfaces(jac, domain, th, t, interfaces)
        yc = y.astype(complex)
        yc[Ns] += 1j * cstep
        jac[:, Ns] = dydtreactor(domain, yc, t, p, interfaces).imag / cstep
    else:  # ConstantTV / ConstantTAPhi
        jacobianynsderiv(jac, domain, th, None)
        for ind in domain.constantspeciesinds:
            jac[ind, :] = 0.0
        _jac_interfaces(jac, domain, th, t, interfaces)
    return jac


def jacobiany_complexstep(domain, y, t=0.0, p=None, h=1e-30, interfaces=()):
    """Stand-in for jacobianyforwarddiff (src/Reactor.jl:771-782): derivative of the full RHS."""
    y = np.asarray(y, dtype=float)
    J = np.zeros((len(y), len(y)))
    for k in range(len(y)):
        yc = y.astype(complex)
        yc[k] += 1j * h
        J[:, k] = dydtreactor(domain, yc, t, p, interfaces).imag / h
    return J


# --------------------------------------------------------------------------------------
# Parameter Jacobian: src/Domain.jl:3095-3485
# --------------------------------------------------------------------------------------
def jacobianp(domain, y, t=0.0, p=None, interfaces=()):
    """d(dy/dt)/dp, Ny x (Ns+Nr).  Columns 0..Ns-1: dG_i (enters through krev = kf/Kc);
    columns Ns+j: k_j (constant-k domains: the rate coefficient itself, or the multiplier for
    efficiency reactions; variable-k domains: the multiplier).  src/Domain.jl:3095-3374.
    Boundary interfaces: the reference adds a term only for an Inlet on ConstantV / ConstantP (src/Domain.jl:3421-3430,3461-3470,
    restated as written there); ConstantTP / TV / TAPhi ignore them (:3376-3392,3473-3485)."""
    y = np.asarray(y, dtype=float)
    ph = domain.phase
    Ns, Nr = ph.Ns, ph.Nr
    Ny = len(y)
    th = domain.calcthermo(y, t, p)
    jacp = np.zeros((Ny, Ns + Nr))
    RTinv = 1.0 / (R * th.T)
    ra = ph.rxnarray
    constk = domain.kind in ("ConstantTP", "ConstantTAPhi", "ConstantTV")
    for i in range(Nr):
        rs = [ra[j, i] for j in range(4) if ra[j, i] >= 0]
        ps = [ra[j, i] for j in range(4, 8) if ra[j, i] >= 0]
        prodr = 1.0
        for s in rs:
            prodr = prodr * th.cs[s]
        prodp = 1.0
        for s in ps:
            prodp = prodp * th.cs[s]
        if domain.kind == "ConstantTV":   # src/Domain.jl:3290-3317 (diffusion-limited chain rule)
            pk = (domain.p if p is None else p)[Ns + i]
            fderiv = th.kfs[i] * th.kfs[i] / (pk * pk) * prodr
            rderiv = th.kfs[i] * th.krevs[i] / (pk * pk) * prodp
            gderiv = rderiv * (pk * pk) / th.kfs[i] * RTinv
        elif constk:                      # src/Domain.jl:3095-3143 (TP, TAPhi)
            if i in ph.efficiencyinds:
                fderiv = th.kfs[i] * prodr
                rderiv = th.krevs[i] * prodp
            else:
                fderiv = prodr
                rderiv = th.krevs[i] / th.kfs[i] * prodp
            gderiv = rderiv * th.kfs[i] * RTinv
        else:                             # src/Domain.jl:3204-3233 (variable-k domains)
            fderiv = th.kfs[i] * prodr
            rderiv = th.krevs[i] * prodp
            gderiv = rderiv * RTinv
        flux = fderiv - rderiv
        for s in rs:
            jacp[s, Ns + i] -= flux
        for s in ps:
            jacp[s, Ns + i] += flux
        # dG columns (krev = kf/Kc, Kc = exp(-dG/RT)): src/Domain.jl:3144-3199
        for s, sign in [(s, -1.0) for s in rs] + [(s, 1.0) for s in ps]:
            for q in rs:
                jacp[q, s] -= -sign * gderiv
            for q in ps:
                jacp[q, s] += -sign * gderiv
    jacp *= th.V
    if domain.kind in ("ConstantTP", "ParametrizedTP"):        # src/Domain.jl:3376-3392
        for i in range(Ns + Nr):
            jacp[Ns, i] = np.sum(jacp[:Ns, i]) * R * th.T / th.P
    elif domain.kind in ("ConstantV", "ParametrizedV"):        # src/Domain.jl:3394-3431
        dydt = np.zeros(Ny)
        addreactionratecontributions(dydt, ra, th.cs, th.kfs, th.krevs)
        dydt *= th.V
        for i in range(Ns):
            jacp[Ns, i] = -(np.dot(th.Us, jacp[:Ns, i]) + dydt[i]) / (th.N * th.Cvave)
            jacp[Ns + 1, i] = np.sum(jacp[:Ns, i]) * R * th.T / th.V + th.P / th.T * jacp[Ns, i]
        for i in range(Ns, Ns + Nr):
            jacp[Ns, i] = -np.dot(th.Us, jacp[:Ns, i]) / (th.N * th.Cvave)
            jacp[Ns + 1, i] = np.sum(jacp[:Ns, i]) * R * th.T / th.V + th.P / th.T * jacp[Ns, i]
    elif domain.kind in ("ConstantP", "ParametrizedP"):        # src/Domain.jl:3433-3471
        dydt = np.zeros(Ny)
        addreactionratecontributions(dydt, ra, th.cs, th.kfs, th.krevs)
        dydt *= th.V
        Cp = th.Cvave + R
        for i in range(Ns):
            jacp[Ns, i] = -(np.dot(th.Hs, jacp[:Ns, i]) + dydt[i]) / (th.N * Cp)
            jacp[Ns + 1, i] = np.sum(jacp[:Ns, i]) * R * th.T / th.P + jacp[Ns, i] * th.V / th.T
        for i in range(Ns, Ns + Nr):
            jacp[Ns, i] = -np.dot(th.Hs, jacp[:Ns, i]) / (th.N * Cp)
            jacp[Ns + 1, i] = np.sum(jacp[:Ns, i]) * R * th.T / th.P + jacp[Ns, i] * th.V / th.T
    for ind in domain.constantspeciesinds:
        jacp[ind, :] = 0.0
    if domain.kind in ("ConstantV", "ConstantP", "ParametrizedV", "ParametrizedP"):
        for inter in interfaces:
            if isinstance(inter, Inlet) and inter.domain is domain:
                flow = _flow(inter.F, t)
                isv = domain.kind in ("ConstantV", "ParametrizedV")
                Cx = th.Cvave if isv else th.Cvave + R
                for i in range(Ns):
                    ddG = flow * (-th.ns[i] / th.N) / (th.N * Cx)
                    jacp[Ns, i] += ddG
                    jacp[Ns + 1, i] += th.P / th.T * ddG if isv else ddG * th.V / th.T
    return jacp
