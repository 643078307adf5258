"""ctypes binding of librmsb200.so (include/rmsb200.h).  Fails loudly when the CUDA library is
missing -- there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("RMSB200_LIB", _HERE.parent / "csrc" / "librmsb200.so"))

# ---- constants mirrored from include/rmsb200.h -------------------------------------------------
RMSB_OK = 0
PHASE_GAS, PHASE_LIQUID, PHASE_SURFACE = 0, 1, 2
DOMAIN_CONSTANT_V, DOMAIN_CONSTANT_TP, DOMAIN_CONSTANT_TV, DOMAIN_CONSTANT_TAPHI, DOMAIN_COUPLED_TP_TAPHI = 0, 1, 2, 3, 4
DOMAIN_CONSTANT_P = 5
(KIN_ARRHENIUS_VEC, KIN_ARRHENIUS, KIN_MULTIARRHENIUS, KIN_PLOG_VEC, KIN_PLOG, KIN_MULTIPLOG,
 KIN_THIRDBODY, KIN_LINDEMANN, KIN_TROE, KIN_CHEBYSHEV_VEC) = range(10)
VISC_NONE, VISC_CONSTANT, VISC_RIEDEL = 0, 1, 2
SOLVE_OK, SOLVE_MAX_STEPS, SOLVE_STEP_TOO_SMALL, SOLVE_NAN = 0, 1, 2, 3

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)
_pu8 = C.POINTER(C.c_uint8)


class MechDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("phase_kind", C.c_int32), ("n_species", C.c_int32), ("n_reactions", C.c_int32),
        ("sitedensity", C.c_double), ("diffusion_limited", C.c_int32), ("nasa_max_ranges", C.c_int32),
        ("nasa_nranges", _pi), ("nasa_ncoef", _pi), ("nasa_tmax", _pd), ("nasa_coefs", _pd),
        ("rxnarray", _pi), ("kin_type", _pi), ("kin_offset", _pi), ("kin_params", _pd),
        ("eff_ptr", _pi), ("eff_species", _pi), ("eff_value", _pd),
        ("reversible", _pu8), ("forwardable", _pu8), ("radius", _pd), ("diff_radius", _pd),
        ("viscosity_kind", C.c_int32), ("viscosity_params", C.c_double * 5),
        ("species_phase", _pi), ("rxn_domain", _pi),
    ]


class Stats(C.Structure):
    _fields_ = [("n_steps", C.c_int32), ("n_rhs", C.c_int32), ("n_jac", C.c_int32), ("n_lu", C.c_int32),
                ("n_newton_fail", C.c_int32), ("n_err_fail", C.c_int32), ("last_order", C.c_int32),
                ("reserved", C.c_int32), ("last_h", C.c_double)]


STATS_DTYPE = np.dtype([("n_steps", "i4"), ("n_rhs", "i4"), ("n_jac", "i4"), ("n_lu", "i4"), ("n_newton_fail", "i4"),
                        ("n_err_fail", "i4"), ("last_order", "i4"), ("reserved", "i4"), ("last_h", "f8")])
assert STATS_DTYPE.itemsize == C.sizeof(Stats)


class SolveOpts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("max_steps", C.c_int32), ("max_order", C.c_int32), ("idt_index", C.c_int32),
                ("first_step", C.c_double), ("warps_per_block", C.c_int32), ("reserved", C.c_int32),
                ("dysave", C.c_void_p)]          # version 2: host buffer [n][nsave][Ny] for du at the saved states (rmsb_solve_batch)


class Extras(C.Structure):       # rmsb_extras
    _fields_ = [("struct_size", C.c_int32), ("interfaces", C.c_void_p), ("if_stride", C.c_int64), ("profile", C.c_void_p),
                ("profile_stride", C.c_int64), ("t", C.c_void_p)]


class RMSBError(RuntimeError):
    pass


_lib = None

EXPORTS = [
    "rmsb_version", "rmsb_last_error", "rmsb_device_count", "rmsb_mech_create", "rmsb_mech_destroy", "rmsb_mech_ny",
    "rmsb_mech_ncond", "rmsb_mech_device", "rmsb_dydt_batch", "rmsb_jac_batch", "rmsb_rates_batch", "rmsb_gibbs_batch",
    "rmsb_solve_batch", "rmsb_dydt_batch_dev", "rmsb_jac_batch_dev", "rmsb_solve_work_bytes", "rmsb_solve_batch_dev",
    "rmsb_kernel_launch_count", "rmsb_fp64_peak_tflops", "rmsb_jacp_batch", "rmsb_jacp_batch_dev", "rmsb_sens_batch",
    "rmsb_sens_work_bytes", "rmsb_sens_batch_dev", "rmsb_kfs_batch", "rmsb_large_info", "rmsb_large_profile",
    "rmsb_solve_batch_multi", "rmsb_last_kernel_ms", "rmsb_dydt_batch_if", "rmsb_jac_batch_if", "rmsb_solve_batch_if",
    "rmsb_dydt_batch_ex", "rmsb_jac_batch_ex", "rmsb_solve_batch_ex", "rmsb_spec_source_desc", "rmsb_mech_specialized",
    "rmsb_jacp_batch_if", "rmsb_sens_batch_if", "rmsb_jacp_batch_ex", "rmsb_sens_batch_ex",
]


def lib():
    """Load (once) and return the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RMSBError(f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.rmsb_version.restype = C.c_int
    L.rmsb_last_error.restype = C.c_char_p
    L.rmsb_device_count.restype = C.c_int
    L.rmsb_mech_create.argtypes = [C.POINTER(MechDesc), C.c_int, C.POINTER(vp)]
    L.rmsb_mech_destroy.argtypes = [vp]
    L.rmsb_mech_destroy.restype = None
    for f in (L.rmsb_mech_ny, L.rmsb_mech_ncond):
        f.argtypes = [vp, C.c_int]
    L.rmsb_mech_device.argtypes = [vp]
    common = [vp, C.c_int, i64, vp, vp, vp, i64]
    L.rmsb_dydt_batch.argtypes = common + [vp, i32, vp]
    L.rmsb_jac_batch.argtypes = common + [vp, i32, vp]
    L.rmsb_rates_batch.argtypes = common + [vp, vp]
    L.rmsb_gibbs_batch.argtypes = [vp, i64, vp, C.c_int, vp]
    L.rmsb_solve_batch.argtypes = common + [vp, i32, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, C.POINTER(SolveOpts)]
    L.rmsb_dydt_batch_dev.argtypes = common + [vp, i32, vp, vp]
    L.rmsb_jac_batch_dev.argtypes = common + [vp, i32, vp, vp]
    L.rmsb_solve_work_bytes.argtypes = [vp, C.c_int, i64]
    L.rmsb_solve_work_bytes.restype = i64
    L.rmsb_solve_batch_dev.argtypes = common + [vp, i32, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, C.POINTER(SolveOpts), vp, vp]
    L.rmsb_large_info.argtypes = [vp, vp]
    L.rmsb_large_profile.argtypes = [vp, C.c_int, vp]
    L.rmsb_kfs_batch.argtypes = [vp, C.c_int, i64, vp, vp, vp]
    L.rmsb_jacp_batch.argtypes = common + [vp, i32, vp]
    L.rmsb_jacp_batch_dev.argtypes = common + [vp, i32, vp, vp]
    L.rmsb_sens_batch.argtypes = common + [vp, i32, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(SolveOpts)]
    L.rmsb_sens_work_bytes.argtypes = [vp, C.c_int, i64]
    L.rmsb_sens_work_bytes.restype = i64
    L.rmsb_sens_batch_dev.argtypes = common + [vp, i32, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(SolveOpts), vp, vp]
    L.rmsb_kernel_launch_count.restype = i64
    L.rmsb_solve_batch_multi.argtypes = [vp, i32, C.c_int, i64, vp, vp, vp, i64, vp, i32, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp,
                                         C.POINTER(SolveOpts)]
    L.rmsb_last_kernel_ms.restype = dbl
    L.rmsb_dydt_batch_if.argtypes = common + [vp, i32, vp, i64, vp]
    L.rmsb_jac_batch_if.argtypes = common + [vp, i32, vp, i64, vp]
    L.rmsb_solve_batch_if.argtypes = common + [vp, i32, vp, i64, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, C.POINTER(SolveOpts)]
    L.rmsb_dydt_batch_ex.argtypes = common + [vp, i32, C.POINTER(Extras), vp]
    L.rmsb_jac_batch_ex.argtypes = common + [vp, i32, C.POINTER(Extras), vp]
    L.rmsb_solve_batch_ex.argtypes = common + [vp, i32, C.POINTER(Extras), vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, C.POINTER(SolveOpts)]
    L.rmsb_jacp_batch_if.argtypes = common + [vp, i32, vp, i64, vp]
    L.rmsb_jacp_batch_ex.argtypes = common + [vp, i32, C.POINTER(Extras), vp]
    L.rmsb_sens_batch_ex.argtypes = common + [vp, i32, C.POINTER(Extras), vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(SolveOpts)]
    L.rmsb_sens_batch_if.argtypes = common + [vp, i32, vp, i64, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(SolveOpts)]
    L.rmsb_spec_source_desc.argtypes = [C.POINTER(MechDesc), C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.rmsb_mech_specialized.argtypes = [vp, C.c_int]
    L.rmsb_fp64_peak_tflops.argtypes = [C.c_int, C.c_int]
    L.rmsb_fp64_peak_tflops.restype = dbl
    _lib = L
    return L


def check(rc: int):
    if rc != RMSB_OK:
        raise RMSBError(f"librmsb200 error {rc}: {lib().rmsb_last_error().decode()}")


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
