"""The thread-per-reactor dydt / Jacobian kernels (batches >= 4096 of small ConstantV / ConstantP mechanisms) against the oracle and
against the kernels that smaller batches use: both the kernels generated and compiled for the mechanism (csrc/rmsb_spec.cu) and the
generic table-driven ones they replace (RMSB_NO_SPEC=1)."""
import os

import numpy as np
import pytest

import cases
import rmsb200 as rb
from oracle import rms_oracle as ro
from rmsb200 import _lib as L

pytestmark = pytest.mark.gpu


def _states(y0, Ns, n, seed):
    rng = np.random.default_rng(seed)
    ys = np.repeat(y0[None], n, axis=0)
    tot = y0[:Ns].sum()
    ys[:, :Ns] = np.abs(ys[:, :Ns] + tot * 10.0 ** rng.uniform(-8, -2, (n, Ns)) * rng.uniform(0.1, 1, (n, Ns)))
    ys[:, Ns] *= rng.uniform(0.9, 1.6, n)
    ys[:, Ns + 1] *= rng.uniform(0.8, 1.5, n)
    return ys


@pytest.fixture(params=["specialised", "generic"])
def big_batch_kernels(request):
    old = os.environ.get("RMSB_NO_SPEC")
    os.environ["RMSB_NO_SPEC"] = "0" if request.param == "specialised" else "1"
    yield request.param
    if old is None:
        os.environ.pop("RMSB_NO_SPEC", None)
    else:
        os.environ["RMSB_NO_SPEC"] = old


@pytest.mark.parametrize("which", ["ConstantV", "ConstantP"])
def test_thread_per_reactor_dydt(which, big_batch_kernels):
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    ic = {"T": 1000.0, "P": 10e5, "H2": 0.67, "O2": 0.33, "N2": 0.5}
    d = rb.readinput(cases.fixture("superminimal"))
    ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
    od = ro.readinput(cases.fixture("superminimal"))
    oig = ro.IdealGas(od["phase"]["Species"], od["phase"]["Reactions"])
    if which == "ConstantV":
        gdom, y0, p = rb.ConstantVDomain(phase=ig, initialconds=ic, constantspecies=["N2"])
        odom = ro.ConstantVDomain(oig, ic, constantspecies=("N2",))
    else:
        gdom, y0, p = rb.ConstantPDomain(phase=ig, initialconds=ic)
        odom = ro.ConstantPDomain(oig, ic)
    # the run-time compiled kernels must really be the ones that run (NVRTC is part of the CUDA toolkit of the image)
    assert ig.specialized(gdom.kind) == (big_batch_kernels == "specialised"), L.lib().rmsb_last_error().decode()
    n = 6000
    ys = _states(y0, ig.Ns, n, 3)
    ens = rb.Domain(gdom.kind, ig, ys, np.repeat(gdom.cond, n, axis=0), constantspecies=["N2"] if which == "ConstantV" else [])
    ens.p, ens._p_default = gdom.p, gdom._p_default
    pp = cases.parameter_vector("superminimal_V", ig.Ns, ig.Nr)
    for par in (p, pp):
        f_big = ens.dydt(ys, par)                                   # n >= 4096: one thread per reactor
        f_small = np.concatenate([ens_chunk(ens, ys[k:k + 1000]).dydt(ys[k:k + 1000], par) for k in range(0, n, 1000)])   # warp per reactor
        scale = np.abs(f_small) + 1e-6 * np.abs(f_small).max(axis=1, keepdims=True) + 1e-300
        # two kernels, two summation orders of the cancelling dot(U, dn/dt) in dT/dt: a few 1e-12 of the row scale on the worst of
        # 6000 random states; the bar against the oracle below stays 1e-12
        assert np.max(np.abs(f_big - f_small) / scale) < 5e-12
        for k in (0, 1, 2, n - 1):
            assert cases.err_dydt(f_big[k], ro.dydtreactor(odom, ys[k], 0.0, par)) < 1e-12
        J_big = ens.jacobiany(ys, par)
        J_small = np.concatenate([ens_chunk(ens, ys[k:k + 1000]).jacobiany(ys[k:k + 1000], par) for k in range(0, n, 1000)])
        col = np.max(np.abs(J_small), axis=1, keepdims=True)
        # two kernels, two summation orders (tree vs sequential) of dot(U, dn/dt), whose terms cancel: the T / P columns of the
        # worst of 6000 random states differ by 1e-11 of the column scale; the bar against the oracle below stays 1e-12
        assert np.max(np.abs(J_big - J_small) / (np.abs(J_small) + 1e-3 * col + 1e-300)) < 1e-10
        for k in (0, 1, 2, 3, 4, n - 1):
            assert cases.err_jac(J_big[k], ro.jacobiany(odom, ys[k], 0.0, par)) < 1e-12


    # per-reactor parameter rows: only the reactor whose row differs changes
    per = np.repeat(pp[None], n, axis=0)
    per[7, ig.Ns + 2] *= 3.0
    f_shared, f_per = ens.dydt(ys, pp), ens.dydt(ys, per)
    keep = np.arange(n) != 7
    assert np.array_equal(f_shared[keep], f_per[keep]) and not np.array_equal(f_shared[7], f_per[7])
    assert cases.err_dydt(f_per[7], ro.dydtreactor(odom, ys[7], 0.0, per[7])) < 1e-12


def ens_chunk(ens, ys):
    sub = rb.Domain(ens.kind, ens.phase, ys, np.repeat(ens.cond[:1], len(ys), axis=0))
    sub.constantspeciesinds = ens.constantspeciesinds
    sub.p, sub._p_default = ens.p, ens._p_default
    return sub
