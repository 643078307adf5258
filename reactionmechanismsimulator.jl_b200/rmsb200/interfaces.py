"""Boundary interfaces (src/Interface.jl:234-365): Inlet, Outlet, VolumetricFlowRateInlet, VolumetricFlowRateOutlet and
VolumeMaintainingOutlet turn the batch reactors into CSTR / PFR-like ensembles.  The constructors mirror the reference's; the flows
must be CONSTANT in time here (a callable is accepted when it returns the same value at several times) -- the device evaluates
them inside dydt / jacobiany / solve (rmsb_*_batch_if, include/rmsb200.h) for every reactor of the ensemble, each with its own row.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib as L
from .mechanism import NASA, R

_SPECIAL = ("T", "P", "Hin")


def get_enthalpy(thermo: NASA, T: float) -> float:
    """getEnthalpy(NASA, T): selectPoly + the scalar polynomial with exact c_k/k (src/Calculators/Thermo.jl:82-93,101-111)."""
    poly = thermo.polys[-1]
    for q in thermo.polys:
        if T <= q.Tmax:
            poly = q
            break
    c = list(poly.coefs)

    def horner(coefs):
        acc = 0.0
        for k in range(len(coefs) - 1, -1, -1):
            acc = acc + coefs[k]
            if k != 0:
                acc = acc * T
        return acc

    if len(c) == 9:
        return ((-c[0] / T + c[1] * math.log(T)) / T + horner([c[2 + k] / (k + 1) for k in range(5)])) * R * T + c[7] * R
    return horner([c[k] / (k + 1) for k in range(5)]) * R * T + c[5] * R


def _constant(F, what: str) -> float:
    if callable(F):
        vals = [float(F(t)) for t in (0.0, 1.0e-3, 1.0, 1.0e3)]
        if any(v != vals[0] for v in vals):
            raise NotImplementedError(f"time-dependent {what}: the accelerated hot path evaluates constant flows (DESIGN.md)")
        return vals[0]
    return float(F)


def _spcsvector(domain, conddict: Dict[str, float]) -> np.ndarray:
    v = np.zeros(domain.phase.Ns)
    for k, val in conddict.items():
        if k not in _SPECIAL:
            if k not in domain.phase.spcdict:
                raise KeyError(f"{k} not found in species list")
            v[domain.phase.spcdict[k] - 1] = float(val)
    return v


class Inlet:
    """Inlet(domain, conddict, F) (src/Interface.jl:243-263): molar flow F of mole fractions y and enthalpy H."""

    def __init__(self, domain, conddict: Dict[str, float], F):
        y = _spcsvector(domain, conddict)
        self.domain, self.y, self.F = domain, y / y.sum(), _constant(F, "Inlet flow")
        self.T, self.P = float(conddict.get("T", 0.0)), float(conddict.get("P", 0.0))
        if "Hin" in conddict:
            self.H = float(conddict["Hin"])
        else:
            if self.T == 0.0 or self.P == 0.0:
                raise ValueError("T and P need to be provided if Hin is not provided for Inlet")
            self.H = float(sum(get_enthalpy(s.thermo, self.T) * self.y[i] for i, s in enumerate(domain.phase.species)))


class Outlet:
    """Outlet(domain, F) (src/Interface.jl:265-269)."""

    def __init__(self, domain, F):
        self.domain, self.F = domain, _constant(F, "Outlet flow")


class VolumetricFlowRateInlet:
    """VolumetricFlowRateInlet(domain, conddict, Vin) (src/Interface.jl:320-353): concentrations cs, enthalpy per volume."""

    def __init__(self, domain, conddict: Dict[str, float], Vin):
        self.domain, self.cs, self.Vin = domain, _spcsvector(domain, conddict), _constant(Vin, "volumetric inlet flow")
        self.T, self.P = float(conddict.get("T", 0.0)), float(conddict.get("P", 0.0))
        if "Hin" in conddict:
            self.Hpervolume = float(conddict["Hin"]) / self.cs.sum()
        else:
            if self.T == 0.0 or self.P == 0.0:
                raise ValueError("T and P need to be provided if Hin is not provided for Inlet")
            self.Hpervolume = float(sum(get_enthalpy(s.thermo, self.T) * self.cs[i] for i, s in enumerate(domain.phase.species)))


class VolumetricFlowRateOutlet:
    """VolumetricFlowRateOutlet(domain, Vout) (src/Interface.jl:314-318)."""

    def __init__(self, domain, Vout):
        self.domain, self.Vout = domain, _constant(Vout, "volumetric outlet flow")


class VolumeMaintainingOutlet:
    """VolumeMaintainingOutlet(domain) (src/Interface.jl:356-365): ConstantTP / ConstantP domains."""

    def __init__(self, domain):
        if domain.kind not in (L.DOMAIN_CONSTANT_TP, L.DOMAIN_CONSTANT_P):
            raise ValueError("VolumeMaintainingOutlet needs a domain whose volume is a state (ConstantTP, ConstantP)")
        self.domain = domain


BOUNDARY_TYPES = (Inlet, Outlet, VolumetricFlowRateInlet, VolumetricFlowRateOutlet, VolumeMaintainingOutlet)


def pack_interfaces(domain, interfaces: Sequence, flows: Optional[np.ndarray] = None) -> np.ndarray:
    """One row [F_in, F_out, Vin, Vout, volume-maintaining, H_in, Hpervolume, y_in[Ns], cs_in[Ns]] (rmsb200.h) for the interfaces
    attached to `domain`.  Several interfaces of one type are merged: the terms are linear in flow*composition and flow*enthalpy.
    `flows` [n][4] optionally gives every reactor of the ensemble its own (F_in, F_out, Vin, Vout): a residence-time sweep."""
    Ns = domain.phase.Ns
    row = np.zeros(7 + 2 * Ns)
    for it in interfaces:
        if not isinstance(it, BOUNDARY_TYPES):
            raise NotImplementedError(f"interface {type(it).__name__} is outside the accelerated hot path (DESIGN.md)")
        if it.domain is not domain:
            continue
        if isinstance(it, Inlet):
            tot = row[0] + it.F
            if tot != 0.0:
                row[7:7 + Ns] = (row[7:7 + Ns] * row[0] + it.y * it.F) / tot
                row[5] = (row[5] * row[0] + it.H * it.F) / tot
            row[0] = tot
        elif isinstance(it, Outlet):
            row[1] += it.F
        elif isinstance(it, VolumetricFlowRateInlet):
            tot = row[2] + it.Vin
            if tot != 0.0:
                row[7 + Ns:] = (row[7 + Ns:] * row[2] + it.cs * it.Vin) / tot
                row[6] = (row[6] * row[2] + it.Hpervolume * it.Vin) / tot
            row[2] = tot
        elif isinstance(it, VolumetricFlowRateOutlet):
            row[3] += it.Vout
        else:
            row[4] = 1.0
    if flows is None:
        return row
    flows = np.atleast_2d(np.asarray(flows, dtype=float))
    out = np.repeat(row[None], flows.shape[0], axis=0)
    out[:, :4] = flows
    return np.ascontiguousarray(out)
