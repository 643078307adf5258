"""ctypes wrapper of oracle/cpu/librms_cpu_oracle.so -- the C++ CPU port (TEST INFRASTRUCTURE / CPU
BASELINE ONLY, see oracle/__init__.py).  Flattens the ORACLE's own Phase objects (oracle/rms_oracle.py),
not the product's, into the descriptor layout of include/rmsb200.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from . import rms_oracle as ro

HERE = Path(__file__).resolve().parent
LIB = HERE / "cpu" / "librms_cpu_oracle.so"
SRC = HERE / "cpu" / "rms_cpu.cpp"

KIN = dict(ARRHENIUS_VEC=0, ARRHENIUS=1, MULTIARRHENIUS=2, PLOG_VEC=3, PLOG=4, MULTIPLOG=5, THIRDBODY=6, LINDEMANN=7, TROE=8, CHEBYSHEV_VEC=9)
DOMAIN = {"ConstantV": 0, "ConstantTP": 1, "ConstantTV": 2, "ConstantTAPhi": 3, "ConstantP": 5}
PHASE = {"gas": 0, "liquid": 1, "surface": 2}

_pd, _pi, _pu8 = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)


class Desc(C.Structure):   # rmsb_mech_desc
    _fields_ = [("struct_size", C.c_int32), ("phase_kind", C.c_int32), ("n_species", C.c_int32), ("n_reactions", C.c_int32),
                ("sitedensity", C.c_double), ("diffusion_limited", C.c_int32), ("nasa_max_ranges", C.c_int32),
                ("nasa_nranges", _pi), ("nasa_ncoef", _pi), ("nasa_tmax", _pd), ("nasa_coefs", _pd),
                ("rxnarray", _pi), ("kin_type", _pi), ("kin_offset", _pi), ("kin_params", _pd),
                ("eff_ptr", _pi), ("eff_species", _pi), ("eff_value", _pd), ("reversible", _pu8), ("forwardable", _pu8),
                ("radius", _pd), ("diff_radius", _pd), ("viscosity_kind", C.c_int32), ("viscosity_params", C.c_double * 5),
                ("species_phase", _pi), ("rxn_domain", _pi)]


STATS = np.dtype([("n_steps", "i4"), ("n_rhs", "i4"), ("n_jac", "i4"), ("n_lu", "i4"), ("n_newton_fail", "i4"),
                  ("n_err_fail", "i4"), ("last_order", "i4"), ("reserved", "i4"), ("last_h", "f8")])


STRICT_FLAGS = ["-O2", "-fno-fast-math", "-ffp-contract=off"]                       # the checker: no FMA contraction, portable code
# the CPU BASELINE build (bench.py only): what a tuned native build of the same source gives on this host -- the reference runs its
# hot path under @fastmath with native code generation.  -ffinite-math-only is left out: the integrator tests isfinite().
FAST_FLAGS = ["-O3", "-march=native", "-ffast-math", "-fno-finite-math-only", "-funroll-loops"]
_build_kind = "strict"


def _fast_lib_path():
    """-march=native code must not travel between hosts: one file per CPU model."""
    import hashlib
    try:
        model = [l for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags"))][:2]
    except OSError:
        model = ["unknown"]
    return HERE / "cpu" / f"librms_cpu_fast_{hashlib.sha1(''.join(model).encode()).hexdigest()[:10]}.so"


def build(force=False, kind=None):
    kind = kind or _build_kind
    out = LIB if kind == "strict" else _fast_lib_path()
    if force or not out.exists() or out.stat().st_mtime < SRC.stat().st_mtime:
        flags = STRICT_FLAGS if kind == "strict" else FAST_FLAGS
        tmp = out.with_suffix(f".so.{os.getpid()}.tmp")
        subprocess.run(["g++", *flags, "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-o", str(tmp), str(SRC)], check=True)
        os.replace(tmp, out)
    return out


def select_build(kind: str):
    """"strict" (default: the checker) or "fast" (CPU baseline timing, bench.py).  Handles made before the switch keep their library."""
    global _build_kind, _lib
    assert kind in ("strict", "fast")
    if kind != _build_kind:
        _build_kind, _lib = kind, None


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.rmso_create.argtypes = [C.POINTER(Desc)]
        L.rmso_create.restype = vp
        L.rmso_destroy.argtypes = [vp]
        L.rmso_eval.argtypes = [vp, C.c_int, C.c_int, i64, vp, vp, vp, i64, vp, i32, vp, C.c_int]
        L.rmso_solve.argtypes = [vp, C.c_int, i64, vp, vp, vp, i64, vp, i32, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, i32, i32, C.c_int]
        _lib = L
    return _lib


def _arr(a):
    return [a["A"], a["n"], a["Ea"]]


def _plog(kin):
    out = [float(len(kin["Ps"]))] + list(kin["Ps"])
    for a in kin["arrs"]:
        out += ([1.0] + _arr(a)) if a["type"] == "Arrhenius" else ([float(len(a["arrs"]))] + [v for q in a["arrs"] for v in _arr(q)])
    return out


def _pack(kin, vec):
    t = kin["type"]
    if t == "Arrhenius":
        return (KIN["ARRHENIUS_VEC"] if vec else KIN["ARRHENIUS"]), _arr(kin)
    if t == "MultiArrhenius":
        return KIN["MULTIARRHENIUS"], [float(len(kin["arrs"]))] + [v for q in kin["arrs"] for v in _arr(q)]
    if t == "PdepArrhenius":
        if vec:
            return KIN["PLOG_VEC"], [float(len(kin["Ps"]))] + list(kin["Ps"]) + [v for q in kin["arrs"] for v in _arr(q)]
        return KIN["PLOG"], _plog(kin)
    if t == "MultiPdepArrhenius":
        return KIN["MULTIPLOG"], [float(len(kin["parrs"]))] + [v for pk in kin["parrs"] for v in _plog(pk)]
    if t == "ThirdBody":
        return KIN["THIRDBODY"], _arr(kin["arr"])
    if t == "Lindemann":
        return KIN["LINDEMANN"], _arr(kin["arrhigh"]) + _arr(kin["arrlow"])
    if t == "Troe":
        return KIN["TROE"], _arr(kin["arrhigh"]) + _arr(kin["arrlow"]) + [kin["a"], kin["T3"], kin["T1"], kin["T2"]]
    if t == "Chebyshev":
        nT, nP = kin["coefs"].shape
        return KIN["CHEBYSHEV_VEC"], [kin["Tmin"], kin["Tmax"], kin["Pmin"], kin["Pmax"], float(nT), float(nP)] + list(kin["coefs"].reshape(-1))
    raise NotImplementedError(t)


class CpuMech:
    """C++ handle for an oracle Phase."""

    def __init__(self, phase: ro.Phase):
        self.phase = phase
        Ns, Nr = phase.Ns, phase.Nr
        maxr = max(len(s.polys) for s in phase.species)
        k = self._keep = {}
        k["nr"] = np.array([len(s.polys) for s in phase.species], dtype=np.int32)
        k["nc"] = np.full((Ns, maxr), 7, dtype=np.int32)
        k["tmax"] = np.zeros((Ns, maxr))
        k["coef"] = np.zeros((Ns, maxr, 9))
        for i, s in enumerate(phase.species):
            for r, p in enumerate(s.polys):
                k["nc"][i, r] = len(p.coefs)
                k["tmax"][i, r] = p.Tmax
                k["coef"][i, r, :len(p.coefs)] = p.coefs
        k["rx"] = np.ascontiguousarray((phase.rxnarray.T + 1).astype(np.int32))
        kt, ko, kp, ep, es, ev = [], [0], [], [0], [], []
        for j, r in enumerate(phase.reactions):
            tag, pars = _pack(r.kinetics, j < phase.nvec)
            kt.append(tag)
            kp += [float(v) for v in pars]
            ko.append(len(kp))
            eff = r.kinetics.get("efficiencies") or {}
            for s in sorted(eff):
                es.append(s + 1)
                ev.append(eff[s])
            ep.append(len(es))
        k["kt"], k["ko"], k["kp"] = np.array(kt, dtype=np.int32), np.array(ko, dtype=np.int32), np.array(kp or [0.0])
        k["ep"], k["es"], k["ev"] = np.array(ep, dtype=np.int32), np.array(es or [0], dtype=np.int32), np.array(ev or [0.0])
        k["rev"] = phase.reversibility.astype(np.uint8)
        k["fwd"] = phase.forwardability.astype(np.uint8)
        k["rad"] = np.array([s.radius for s in phase.species])
        k["dr"] = np.array([s.diff_r if s.diff_r is not None else -1.0 for s in phase.species])
        d = Desc()
        d.struct_size, d.phase_kind, d.n_species, d.n_reactions = C.sizeof(Desc), PHASE[phase.kind], Ns, Nr
        d.sitedensity, d.diffusion_limited, d.nasa_max_ranges = phase.sitedensity, int(phase.diffusionlimited), maxr
        pi, pd, pu = (lambda a: a.ctypes.data_as(_pi)), (lambda a: a.ctypes.data_as(_pd)), (lambda a: a.ctypes.data_as(_pu8))
        d.nasa_nranges, d.nasa_ncoef, d.nasa_tmax, d.nasa_coefs = pi(k["nr"]), pi(k["nc"]), pd(k["tmax"]), pd(k["coef"])
        d.rxnarray, d.kin_type, d.kin_offset, d.kin_params = pi(k["rx"]), pi(k["kt"]), pi(k["ko"]), pd(k["kp"])
        d.eff_ptr, d.eff_species, d.eff_value = pi(k["ep"]), pi(k["es"]), pd(k["ev"])
        d.reversible, d.forwardable, d.radius, d.diff_radius = pu(k["rev"]), pu(k["fwd"]), pd(k["rad"]), pd(k["dr"])
        d.viscosity_kind = 0
        if phase.solvent is not None:
            mu = phase.solvent.mu
            if mu["type"] == "RiedelViscosity":
                d.viscosity_kind = 2
                for q, key in enumerate("ABCDE"):
                    d.viscosity_params[q] = mu[key]
            else:
                d.viscosity_kind, d.viscosity_params[0] = 1, mu["mu"]
        self.desc = d
        self.L = lib()
        self.h = self.L.rmso_create(C.byref(d))

    def __del__(self):
        try:
            self.L.rmso_destroy(self.h)
        except Exception:
            pass

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def _common(self, kind, y, cond, p, constspc):
        y = np.ascontiguousarray(np.atleast_2d(y), dtype=float)
        cond = np.ascontiguousarray(np.atleast_2d(cond), dtype=float)
        if cond.shape[0] == 1 and y.shape[0] > 1:
            cond = np.ascontiguousarray(np.repeat(cond, y.shape[0], axis=0))
        ps = 0
        if p is not None:
            p = np.ascontiguousarray(p, dtype=float)
            ps = 0 if p.ndim == 1 else p.shape[1]
        cs = np.array([c + 1 for c in constspc], dtype=np.int32) if len(constspc) else None
        return y, cond, p, ps, cs

    def dydt(self, kind, y, cond, p=None, constspc=(), nthreads=0):
        y, cond, p, ps, cs = self._common(kind, y, cond, p, constspc)
        out = np.empty_like(y)
        rc = self.L.rmso_eval(self.h, 0, DOMAIN[kind], y.shape[0], self._p(y), self._p(cond), self._p(p), ps, self._p(cs),
                             0 if cs is None else cs.size, self._p(out), nthreads)
        assert rc == 0
        return out

    def jac(self, kind, y, cond, p=None, constspc=(), nthreads=0):
        y, cond, p, ps, cs = self._common(kind, y, cond, p, constspc)
        n, Ny = y.shape
        out = np.empty((n, Ny * Ny))
        rc = self.L.rmso_eval(self.h, 1, DOMAIN[kind], n, self._p(y), self._p(cond), self._p(p), ps, self._p(cs),
                             0 if cs is None else cs.size, self._p(out), nthreads)
        assert rc == 0
        return out.reshape(n, Ny, Ny).transpose(0, 2, 1)

    def solve(self, kind, y0, cond, t_end, rtol, atol, p=None, constspc=(), tsave=None, max_steps=500000, idt_index=-1, nthreads=0, flags=0):
        self.L.rmso_set_flags(int(flags))     # same meaning as RMSB200(flags=...): 0 = the CUDA path's defaults, 3 = scipy.integrate.BDF's step logic
        y0, cond, p, ps, cs = self._common(kind, y0, cond, p, constspc)
        n, Ny = y0.shape
        te = np.ascontiguousarray(np.broadcast_to(np.asarray(t_end, dtype=float), (n,)))
        ts = None if tsave is None else np.ascontiguousarray(tsave, dtype=float)
        ns = 0 if ts is None else ts.size
        ysave = np.empty((n, ns, Ny)) if ns else None
        yend, idt = np.empty((n, Ny)), np.empty(n)
        status, stats = np.empty(n, dtype=np.int32), np.zeros(n, dtype=STATS)
        rc = self.L.rmso_solve(self.h, DOMAIN[kind], n, self._p(y0), self._p(cond), self._p(p), ps, self._p(cs),
                              0 if cs is None else cs.size, self._p(te), rtol, atol, ns, self._p(ts), self._p(ysave), self._p(yend),
                              self._p(idt), self._p(status), self._p(stats), max_steps, idt_index, nthreads)
        assert rc == 0
        return dict(ysave=ysave, yend=yend, idt=idt, status=status, stats=stats)
