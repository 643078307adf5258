"""C-ABI edge cases through ctypes on a GPU box: empty batches, bad arguments, per-reactor failure codes instead of global
aborts, NaN inputs, shared / per-reactor parameter strides -- the error behaviour the header promises (include/rmsb200.h)."""
import ctypes as C

import numpy as np
import pytest

import cases
import rmsb200 as rb
from rmsb200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dom():
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    return cases.product_domain("superminimal_V")


def test_empty_batch_is_a_no_op(dom):
    gdom, y0, p = dom
    lib = L.lib()
    assert lib.rmsb_dydt_batch(gdom.phase.handle, gdom.kind, 0, None, None, None, 0, None, 0, None) == 0
    assert lib.rmsb_solve_batch(gdom.phase.handle, gdom.kind, 0, None, None, None, 0, None, 0, None, 1e-6, 1e-16, 0, None, None,
                                None, None, None, None, None) == 0


def test_bad_arguments_return_codes_and_messages(dom):
    gdom, y0, p = dom
    lib = L.lib()
    y = np.ascontiguousarray(gdom.y0)
    out = np.empty_like(y)
    assert lib.rmsb_dydt_batch(None, gdom.kind, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, L.ptr(out)) == -1
    assert b"null mechanism" in lib.rmsb_last_error()
    assert lib.rmsb_dydt_batch(gdom.phase.handle, 77, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, L.ptr(out)) == -1
    assert b"domain kind" in lib.rmsb_last_error()
    assert lib.rmsb_dydt_batch(gdom.phase.handle, gdom.kind, 1, None, L.ptr(gdom.cond), None, 0, None, 0, L.ptr(out)) == -1
    assert lib.rmsb_dydt_batch(gdom.phase.handle, gdom.kind, 1, L.ptr(y), L.ptr(gdom.cond), L.ptr(np.ones(29)), 7, None, 0, L.ptr(out)) == -1
    assert b"p_stride" in lib.rmsb_last_error()
    # a gas mechanism cannot be a surface domain, and the coupled domain needs a two-phase descriptor
    assert lib.rmsb_dydt_batch(gdom.phase.handle, L.DOMAIN_CONSTANT_TAPHI, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, L.ptr(out)) == -1
    assert lib.rmsb_dydt_batch(gdom.phase.handle, L.DOMAIN_COUPLED_TP_TAPHI, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, L.ptr(out)) == -1
    te, st = np.array([0.1]), np.zeros(1, dtype=np.int32)
    assert lib.rmsb_solve_batch(gdom.phase.handle, gdom.kind, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, L.ptr(te), -1.0, 1e-16,
                                0, None, None, None, None, L.ptr(st), None, None) == -1            # rtol must be positive
    with pytest.raises(rb.RMSBError):
        L.check(-3)


def test_failures_are_per_reactor(dom):
    """One reactor with a NaN state and one that runs out of steps do not disturb their neighbours."""
    gdom, y0, p = dom
    ys = np.repeat(gdom.y0, 4, axis=0)
    ys[1, 4] = np.nan
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, 4, axis=0))
    ens.p, ens._p_default = gdom.p, gdom._p_default
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.05)), abstol=1e-16, reltol=1e-6)
    assert list(sol.status) == [L.SOLVE_OK, 3, L.SOLVE_OK, L.SOLVE_OK] and sol.retcode[1] == "Unstable"
    assert np.array_equal(sol.yend[0], sol.yend[2]) and np.isfinite(sol.yend[[0, 2, 3]]).all()
    few = rb.solve(rb.Reactor(ens, None, (0.0, 0.05)), rb.RMSB200(max_steps=20), abstol=1e-16, reltol=1e-6)
    assert list(few.status) == [1, 3, 1, 1] and few.retcode[0] == "MaxIters" and np.all(few.stats["n_steps"][[0, 2, 3]] == 20)


def test_parameter_strides(dom):
    gdom, y0, p = dom
    ys = cases.states("superminimal_V", gdom.y0[0], gdom.phase.Ns, "ConstantV")
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.p, ens._p_default = gdom.p, gdom._p_default
    pp = cases.parameter_vector("superminimal_V", gdom.phase.Ns, gdom.phase.Nr)
    per = np.repeat(pp[None], len(ys), axis=0)
    per[3, gdom.phase.Ns + 2] *= 3.0
    f_shared, f_per = ens.dydt(ys, pp), ens.dydt(ys, per)
    assert np.array_equal(f_shared[[0, 1, 2, 4, 5]], f_per[[0, 1, 2, 4, 5]]) and not np.array_equal(f_shared[3], f_per[3])
    s1 = rb.solve(rb.Reactor(ens, None, (0.0, 1e-3), p=per), abstol=1e-16, reltol=1e-8)
    one = rb.Domain(gdom.kind, gdom.phase, ys[3:4], gdom.cond)
    s2 = rb.solve(rb.Reactor(one, None, (0.0, 1e-3), p=per[3]), abstol=1e-16, reltol=1e-8)
    assert np.array_equal(s1.yend[3], s2.yend[0])


def test_result_does_not_depend_on_launch_geometry(dom):
    """warps_per_block = 1, 2, 4 or the library's own choice (most resident warps per SM): bit-identical trajectories and
    integrator statistics -- a reactor's arithmetic never depends on which warp or block picks it up from the work queue."""
    gdom, y0, p = dom
    rng = np.random.default_rng(5)
    n = 300
    ys = np.repeat(gdom.y0, n, axis=0)
    cond = np.repeat(gdom.cond, n, axis=0)
    ys[:, -2] = rng.uniform(1000.0, 1400.0, n)            # T0; P0 follows from the fixed volume
    ys[:, -1] = ys[:, :-2].sum(axis=1) * rb.R * ys[:, -2] / cond[:, 0]
    ens = rb.Domain(gdom.kind, gdom.phase, ys, cond)
    ens.p, ens._p_default = gdom.p, gdom._p_default
    ref = None
    for wpb in (0, 1, 2, 4):
        sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.02)), rb.RMSB200(warps_per_block=wpb), abstol=1e-16, reltol=1e-6, saveat=[0.01])
        assert (sol.status == L.SOLVE_OK).all()
        if ref is None:
            ref = sol
            continue
        assert np.array_equal(sol.yend, ref.yend) and np.array_equal(sol.u, ref.u)
        assert np.array_equal(sol.stats["n_steps"], ref.stats["n_steps"]) and np.array_equal(sol.stats["n_rhs"], ref.stats["n_rhs"])


def test_round2_entry_points_reject_what_they_cannot_do(dom):
    """Extended entry points (interfaces / profiles / generated kernels): bad strides, profiles on layouts without one, kernels of a
    mechanism outside the generated class."""
    gdom, y0, p = dom
    lib = L.lib()
    Ns, Nr = gdom.phase.Ns, gdom.phase.Nr
    y = np.ascontiguousarray(gdom.y0)
    Jp = np.empty((1, (Ns + Nr) * gdom.Ny))
    row = np.zeros(7 + 2 * Ns)
    assert lib.rmsb_jacp_batch_if(gdom.phase.handle, gdom.kind, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, L.ptr(row), 5, L.ptr(Jp)) == -1
    assert b"if_stride" in lib.rmsb_last_error()
    assert lib.rmsb_jacp_batch_if(gdom.phase.handle, gdom.kind, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, L.ptr(row), 0, L.ptr(Jp)) == 0
    plain = np.empty_like(Jp)
    assert lib.rmsb_jacp_batch(gdom.phase.handle, gdom.kind, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, L.ptr(plain)) == 0
    assert np.array_equal(Jp, plain)                                    # an all-zero interface row changes nothing
    # a time profile on the surface layout is refused by name; a malformed table too
    x = L.Extras()
    x.struct_size = C.sizeof(L.Extras)
    tab = rb.spline_table([0.0, 1.0, 2.0], [1.0, 2.0, 3.0])
    x.profile, x.profile_stride = tab.ctypes.data, 0
    out = np.empty_like(y)
    assert lib.rmsb_dydt_batch_ex(gdom.phase.handle, L.DOMAIN_CONSTANT_TAPHI, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, C.byref(x), L.ptr(out)) != 0
    bad = tab.copy(); bad[2] = bad[1]                                  # knot times must increase
    x.profile = bad.ctypes.data
    assert lib.rmsb_dydt_batch_ex(gdom.phase.handle, gdom.kind, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, C.byref(x), L.ptr(out)) == -1
    assert b"knot times" in lib.rmsb_last_error()
    x.struct_size = 8
    assert lib.rmsb_dydt_batch_ex(gdom.phase.handle, gdom.kind, 1, L.ptr(y), L.ptr(gdom.cond), None, 0, None, 0, C.byref(x), L.ptr(out)) == -1
    assert b"rmsb_extras size" in lib.rmsb_last_error()
    # generated kernels: only ConstantV / ConstantP of an all-Arrhenius gas mechanism; everything else runs on the generic kernels
    assert lib.rmsb_mech_specialized(gdom.phase.handle, L.DOMAIN_CONSTANT_V) == 1
    assert lib.rmsb_mech_specialized(gdom.phase.handle, L.DOMAIN_CONSTANT_TP) == 0 and b"generic kernels" in lib.rmsb_last_error()
    edom, _, _ = cases.product_domain("ethane_V")                      # Troe / PLOG / Chebyshev kinetics
    assert lib.rmsb_mech_specialized(edom.phase.handle, L.DOMAIN_CONSTANT_V) == 0 and b"does not qualify" in lib.rmsb_last_error()
    n = 4608
    ys = np.repeat(edom.y0, n, axis=0)
    f = rb.Domain(edom.kind, edom.phase, ys, np.repeat(edom.cond, n, axis=0)).dydt(ys)       # large batch: the generic thread-per-reactor kernel
    f1 = edom.dydt(edom.y0)
    assert np.allclose(f[0], f1[0], rtol=1e-11, atol=1e-11 * np.abs(f1).max()) and np.array_equal(f[0], f[-1])
