"""The wider warp-per-reactor kernels (2, 3 and 4 state rows per lane: 33-128 state variables) against the C++ CPU port of the
same integrator: propane ConstantV (Ny = 84, falloff / PLOG kinetics, R = 3), a synthetic 100-species ConstantTV mechanism
(R = 4, the last size before the block-per-reactor path takes over) and ethane ConstantTP with a constant species (R = 1 for
reference)."""
import numpy as np
import pytest

import cases
import rmsb200 as rb
import synth
from oracle import cpu_oracle as co
from oracle import rms_oracle as ro
from rmsb200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"


def _compare(sol, ref, tol):
    assert sol.retcode == ["Success"] and ref["status"][0] == 0
    yr = ref["yend"][0]
    big = np.abs(yr) > 1e-9 * np.abs(yr).max()
    assert np.max(np.abs(sol.yend[0] - yr)[big] / np.abs(yr)[big]) < tol
    assert abs(int(sol.stats["n_steps"][0]) - int(ref["stats"]["n_steps"][0])) <= 0.05 * ref["stats"]["n_steps"][0] + 5


def test_propane_constant_v_three_rows_per_lane():
    gdom, y0, p = cases.product_domain("propane_V")
    odom = cases.oracle_domain("propane_V")
    assert gdom.Ny == 84
    tend = 2.0e-3
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p), abstol=1e-18, reltol=1e-8)
    ref = co.CpuMech(odom.phase).solve("ConstantV", odom.y0[None], odom.cond[None], tend, 1e-8, 1e-18, p=odom.p, nthreads=1)
    _compare(sol, ref, 1e-5)


def test_synthetic_100_species_four_rows_per_lane(tmp_path):
    ns = 100
    path = str(tmp_path / "synth100.rms")
    synth.write_synthetic_mechanism(path, ns=ns, seed=5)
    d = rb.readinput(path)
    ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase", diffusionlimited=False)
    od = ro.readinput(path)
    oph = ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0], diffusionlimited=False)
    ic = {"T": 650.0, "V": 1.0, **synth.initial_moles(ns)}
    gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds=ic)
    odom = ro.ConstantTVDomain(oph, ic)
    assert gdom.Ny == 100
    ys = np.stack([odom.y0, np.abs(odom.y0 * 1.3 + 1e-3)])
    ens = rb.Domain(gdom.kind, ph, ys, np.repeat(gdom.cond, 2, axis=0))
    f, J = ens.dydt(ys), ens.jacobiany(ys)
    for k in range(2):
        assert cases.err_dydt(f[k], ro.dydtreactor(odom, ys[k], 0.0, None)) < 1e-12
        assert cases.err_jac(J[k], ro.jacobiany(odom, ys[k], 0.0, None)) < 1e-12
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 1.0), p=p), abstol=1e-14, reltol=1e-8)
    ref = co.CpuMech(oph).solve("ConstantTV", odom.y0[None], odom.cond[None], 1.0, 1e-8, 1e-14, nthreads=1)
    _compare(sol, ref, 1e-5)


def test_ethane_constant_tp_with_constant_species():
    gdom, y0, p = cases.product_domain("ethane_TP")
    odom = cases.oracle_domain("ethane_TP")
    tend = 5.0e-3
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p), abstol=1e-18, reltol=1e-8)
    ref = co.CpuMech(odom.phase).solve("ConstantTP", odom.y0[None], odom.cond[None], tend, 1e-8, 1e-18, p=odom.p,
                                       constspc=odom.constantspeciesinds, nthreads=1)
    _compare(sol, ref, 1e-5)
    ar = gdom.phase.names.index("Ar")
    assert sol.yend[0, ar] == y0[ar]
