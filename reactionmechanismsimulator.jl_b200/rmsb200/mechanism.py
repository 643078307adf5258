"""Host-side data model and `.rms` reader: the `readinput` / `Species` / `ElementaryReaction`
surface of the reference (src/Parse.jl:253-429, src/Species.jl, src/Reaction.jl), kept so that
reference users can switch without changing their scripts.  Pure host code, runs once per
mechanism; nothing here is on the hot path.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import yaml

R = 8.314          # src/Constants.jl:1  (8.314 exactly)
Na = 6.022e23      # src/Constants.jl:3
kB = R / Na


@dataclass
class NASApolynomial:          # src/Calculators/Thermo.jl:10-14
    coefs: List[float]
    Tmin: float
    Tmax: float


@dataclass
class NASA:                    # src/Calculators/Thermo.jl:95-98
    polys: List[NASApolynomial]


@dataclass
class Species:                 # src/Species.jl:8-26
    name: str
    index: int                 # 1-based, as in the reference
    thermo: NASA
    smiles: str = ""
    radius: float = 0.0
    diffusion_r: Optional[float] = None   # StokesDiffusivity.r (None = EmptyDiffusivity)
    molecularweight: float = 0.0
    atomnums: Optional[Dict[str, int]] = None
    bondnum: int = -1


@dataclass
class ElementaryReaction:      # src/Reaction.jl:6-23
    index: int
    reactants: List[Species]
    products: List[Species]
    kinetics: dict
    reversible: bool = True
    forwardable: bool = True
    electronchange: int = 0

    def __str__(self):         # getrxnstr, src/Reaction.jl:31
        return "+".join(s.name for s in self.reactants) + "<=>" + "+".join(s.name for s in self.products)


@dataclass
class Solvent:                 # src/Solvent.jl
    name: str
    mu: dict


# -- molecular information (reference: RDKit at parse time, src/Parse.jl:127-148) -----------------
_MCGOWAN_ML = {"H": 8.71, "He": 6.75, "Li": 22.23, "C": 16.35, "N": 14.39, "O": 12.43, "F": 10.47, "Ne": 8.51,
               "Si": 26.83, "P": 24.87, "S": 22.91, "Cl": 20.95, "Ar": 18.99, "Br": 26.21, "B": 18.31}   # src/Parse.jl:49-57
_WEIGHT = {"H": 1.008, "C": 12.011, "N": 14.007, "O": 15.999, "Ar": 39.948, "He": 4.003, "Ne": 20.180, "S": 32.067,
           "Cl": 35.453, "F": 18.998}
_VAL = {"C": 4, "N": 3, "O": 2, "H": 1, "S": 2, "F": 1, "Cl": 1, "Br": 1, "B": 3, "P": 3}
_KNOWN = {"H", "Li", "C", "O", "N", "Cl", "S", "Ar", "Ne", "He", "P", "F", "Br", "I", "B"}   # src/Parse.jl:41-42
_BOND = {"-": 1.0, "=": 2.0, "#": 3.0, "~": 1.0, ":": 1.5}


def molecule_counts(smiles: str) -> Tuple[Dict[str, int], int, float]:
    """(atom counts incl. hydrogens, number of bonds, molecular weight kg/mol) of a SMILES string --
    what getatomdictfromrdkit(AddHs(MolFromSmiles(s))) returns in the reference.  Raises KeyError for
    elements the reference's elementdict does not know (e.g. Pt), like the reference does."""
    atoms: List[List] = []
    bonds = 0
    branch: List[Optional[int]] = []
    last: Optional[int] = None
    pending = 1.0
    open_rings: Dict[str, Tuple[int, float]] = {}
    pos = 0
    while pos < len(smiles):
        c = smiles[pos]
        new_atom = None
        if c == "[":
            close = smiles.index("]", pos)
            mt = re.match(r"\d*([A-Z][a-z]?)(H\d*)?", smiles[pos + 1:close])
            hcount = 0
            if mt.group(2):
                hcount = int(mt.group(2)[1:]) if len(mt.group(2)) > 1 else 1
            new_atom = [mt.group(1), hcount, 0.0]
            pos = close + 1
        elif c.isalpha():
            sym = smiles[pos:pos + 2] if smiles[pos:pos + 2] in ("Cl", "Br") else c
            new_atom = [sym, None, 0.0]
            pos += len(sym)
        elif c in _BOND:
            pending = _BOND[c]
            pos += 1
        elif c == "(":
            branch.append(last)
            pos += 1
        elif c == ")":
            last = branch.pop()
            pos += 1
        elif c == ".":
            last = None
            pos += 1
        elif c.isdigit():
            if c in open_rings:
                other, b = open_rings.pop(c)
                b = max(b, pending)
                atoms[other][2] += b
                atoms[last][2] += b
                bonds += 1
            else:
                open_rings[c] = (last, pending)
            pending = 1.0
            pos += 1
        else:
            raise ValueError(f"unsupported SMILES token {c!r} in {smiles!r}")
        if new_atom is not None:
            atoms.append(new_atom)
            me = len(atoms) - 1
            if last is not None:
                atoms[last][2] += pending
                atoms[me][2] += pending
                bonds += 1
            last, pending = me, 1.0
    counts: Dict[str, int] = {}
    for sym, hcount, bsum in atoms:
        if sym not in _KNOWN:
            raise KeyError(sym)
        if hcount is None:
            hcount = max(0, int(round(_VAL.get(sym, 0) - bsum)))
        counts[sym] = counts.get(sym, 0) + 1
        if hcount:
            counts["H"] = counts.get("H", 0) + hcount
            bonds += hcount
    return counts, bonds, sum(_WEIGHT[e] * k for e, k in counts.items()) / 1000.0


def getspeciesradius(atomdict: Dict[str, int], nbonds: int) -> float:
    """McGowan radius (src/Parse.jl:191-205)."""
    vtot = sum(n * _MCGOWAN_ML[a] * 1e-6 for a, n in atomdict.items()) - nbonds * 6.56e-6
    return (0.75 / math.pi * (vtot / Na)) ** (1.0 / 3.0)


# -- kinetics dictionaries ----------------------------------------------------------------------
def _kin(k: dict, index_of: Dict[str, int]) -> dict:
    t = k["type"]
    if t in ("Arrhenius", "StickingCoefficient"):
        return {"type": t, "A": float(k["A"]), "n": float(k["n"]), "Ea": float(k["Ea"])}
    if t == "MultiArrhenius":
        return {"type": t, "arrs": [_kin(a, index_of) for a in k["arrs"]]}
    if t == "PdepArrhenius":
        return {"type": t, "Ps": sorted(float(p) for p in k["Ps"]), "arrs": [_kin(a, index_of) for a in k["arrs"]]}
    if t == "MultiPdepArrhenius":
        return {"type": t, "parrs": [_kin(a, index_of) for a in k["parrs"]]}
    if t in ("ThirdBody", "Lindemann", "Troe"):
        out = {"type": t}
        if t == "ThirdBody":
            out["arr"] = _kin(k["arr"], index_of)
        else:
            out["arrhigh"] = _kin(k["arrhigh"], index_of)
            out["arrlow"] = _kin(k["arrlow"], index_of)
        if t == "Troe":
            out.update(a=float(k["a"]), T3=float(k["T3"]), T1=float(k["T1"]), T2=float(k.get("T2", 0.0)))
        # efficiencies are stored by 1-based species index as (eff - 1): src/Parse.jl:411-413
        out["efficiencies"] = {index_of[nm]: float(v) - 1.0 for nm, v in (k.get("efficiencies") or {}).items()}
        return out
    if t == "Chebyshev":
        return {"type": t, "coefs": np.array(k["coefs"], dtype=float), "Tmin": float(k["Tmin"]), "Tmax": float(k["Tmax"]),
                "Pmin": float(k["Pmin"]), "Pmax": float(k["Pmax"])}
    raise NotImplementedError(f"kinetics type {t!r} is outside the accelerated hot path (see DESIGN.md)")


def readinput(fname: str) -> dict:
    """readinput(fname) for `.rms` YAML mechanisms (src/Parse.jl:253-275,281-429).  Returns
    ``{phase: {"Species": [...], "Reactions": [...]}, frozenset({p1, p2}): {"Reactions": [...]}, "Solvents": [...]}``.
    Chemkin / Cantera conversion (src/yml.jl) needs RMG-Py and is out of scope."""
    ext = fname.rsplit(".", 1)[-1]
    if ext not in ("rms", "rmg", "yml", "yaml"):
        raise ValueError(f"file extension {ext} not understood, use .rms yaml mechanisms")
    with open(fname) as fh:
        doc = yaml.load(fh, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))
    out: dict = {}
    if doc.get("Solvents"):
        out["Solvents"] = [Solvent(s["name"], {k: (v if k == "type" else float(v)) for k, v in s["mu"].items()})
                           for s in doc["Solvents"]]
    where: Dict[str, Tuple[Species, str]] = {}
    for ph in doc["Phases"]:
        spcs = []
        for i, d in enumerate(ph["Species"]):
            if d["thermo"]["type"] != "NASA":
                raise NotImplementedError(f"thermo type {d['thermo']['type']!r} is outside the accelerated hot path")
            nasa = NASA([NASApolynomial([float(c) for c in q["coefs"]], float(q["Tmin"]), float(q["Tmax"]))
                         for q in d["thermo"]["polys"]])
            sp = Species(d["name"], i + 1, nasa, smiles=d.get("smiles", ""), radius=float(d.get("radius", 0.0)),
                         molecularweight=float(d.get("molecularweight", 0.0)))
            if sp.smiles and not all(k in d for k in ("atomnums", "bondnum", "molecularweight")):
                try:
                    sp.atomnums, sp.bondnum, sp.molecularweight = molecule_counts(sp.smiles)
                except (KeyError, ValueError):
                    pass   # the reference warns and carries on (src/Parse.jl:321-326)
            if "diffusion" in d:
                if d["diffusion"]["type"] == "StokesDiffusivity":
                    sp.diffusion_r = float(d["diffusion"]["r"])
            elif sp.atomnums is not None and sp.bondnum >= 0:
                try:
                    sp.diffusion_r = getspeciesradius(sp.atomnums, sp.bondnum)
                    if "radius" not in d:
                        sp.radius = sp.diffusion_r
                except KeyError:
                    pass
            if sp.name in where:
                raise ValueError("Two species in the file have the same name")
            where[sp.name] = (sp, ph["name"])
            spcs.append(sp)
        out[ph["name"]] = {"Species": spcs, "Reactions": []}
    index_of = {nm: sp.index for nm, (sp, _) in where.items()}
    for i, rx in enumerate(doc["Reactions"]):
        phases = []
        for nm in list(rx["reactants"]) + list(rx["products"]):
            if where[nm][1] not in phases:
                phases.append(where[nm][1])
        r = ElementaryReaction(i + 1, [where[n][0] for n in rx["reactants"]], [where[n][0] for n in rx["products"]],
                               _kin(rx["kinetics"], index_of), bool(rx.get("reversible", True)),
                               bool(rx.get("forwardable", True)), int(rx.get("electronchange", 0)))
        key = phases[0] if len(phases) == 1 else frozenset(phases)
        out.setdefault(key, {"Reactions": []})["Reactions"].append(r)
    return out


# -- writing mechanisms back (src/yml.jl:14-16: writeyml(D; path) = YAML.write_file(path, D)) --------------------------
def _kin_dict(k: dict, name_of: Dict[int, str]) -> dict:
    """Inverse of _kin: kinetics as the .rms file spells them (efficiencies by species name, as eff rather than eff - 1)."""
    t = k["type"]
    if t in ("Arrhenius", "StickingCoefficient"):
        return {"type": t, "A": float(k["A"]), "n": float(k["n"]), "Ea": float(k["Ea"])}
    if t == "MultiArrhenius":
        return {"type": t, "arrs": [_kin_dict(a, name_of) for a in k["arrs"]]}
    if t == "PdepArrhenius":
        return {"type": t, "Ps": [float(p) for p in k["Ps"]], "arrs": [_kin_dict(a, name_of) for a in k["arrs"]]}
    if t == "MultiPdepArrhenius":
        return {"type": t, "parrs": [_kin_dict(a, name_of) for a in k["parrs"]]}
    if t in ("ThirdBody", "Lindemann", "Troe"):
        out = {"type": t}
        if t == "ThirdBody":
            out["arr"] = _kin_dict(k["arr"], name_of)
        else:
            out["arrhigh"], out["arrlow"] = _kin_dict(k["arrhigh"], name_of), _kin_dict(k["arrlow"], name_of)
        if t == "Troe":
            out.update(a=float(k["a"]), T3=float(k["T3"]), T1=float(k["T1"]), T2=float(k["T2"]))
        out["efficiencies"] = {name_of[i]: float(v) + 1.0 for i, v in k["efficiencies"].items()}
        return out
    if t == "Chebyshev":
        return {"type": t, "coefs": [[float(v) for v in row] for row in np.asarray(k["coefs"])], "Tmin": float(k["Tmin"]), "Tmax": float(k["Tmax"]),
                "Pmin": float(k["Pmin"]), "Pmax": float(k["Pmax"])}
    raise NotImplementedError(t)


def getmechdict(parsed: dict) -> dict:
    """The mechanism dictionary of a `.rms` file ({"Units", "Phases", "Reactions", "Solvents"}) from what `readinput` returned --
    the D that the reference hands to writeyml (src/yml.jl:14-36 builds the same layout from Chemkin objects)."""
    D: dict = {"Units": {}, "Phases": [], "Reactions": []}
    rxns: List[ElementaryReaction] = []
    for key, val in parsed.items():
        if key == "Solvents":
            D["Solvents"] = [{"name": s.name, "type": "Solvent", "mu": dict(s.mu)} for s in val]
            continue
        rxns += val["Reactions"]
        if isinstance(key, frozenset):
            continue
        phase = {"name": key, "Species": []}
        for sp in val["Species"]:
            d = {"name": sp.name, "type": "Species", "radicalelectrons": 0,
                 "thermo": {"type": "NASA", "polys": [{"type": "NASApolynomial", "coefs": [float(c) for c in q.coefs], "Tmin": float(q.Tmin),
                                                      "Tmax": float(q.Tmax)} for q in sp.thermo.polys]}}
            if sp.smiles:
                d["smiles"] = sp.smiles
            elif sp.radius:
                d["radius"] = float(sp.radius)
            phase["Species"].append(d)
        D["Phases"].append(phase)
    for r in sorted(rxns, key=lambda x: x.index):
        name_of = {s.index: s.name for s in r.reactants + r.products}
        # efficiencies may name species outside the reaction: resolve through every phase
        for val in parsed.values():
            if isinstance(val, dict) and "Species" in val:
                name_of.update({s.index: s.name for s in val["Species"]})
        D["Reactions"].append({"type": "ElementaryReaction", "reactants": [s.name for s in r.reactants], "products": [s.name for s in r.products],
                               "kinetics": _kin_dict(r.kinetics, name_of), "reversible": bool(r.reversible), "forwardable": bool(r.forwardable),
                               "electronchange": int(r.electronchange), "radicalchange": 0})
    return D


def writeyml(D: dict, path: str = "chem.rms") -> None:
    """writeyml(D; path="chem.rms")  src/yml.jl:14-16"""
    with open(path, "w") as fh:
        yaml.dump(D, fh, Dumper=getattr(yaml, "CSafeDumper", yaml.SafeDumper), default_flow_style=None, sort_keys=True)
