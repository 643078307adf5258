"""Config C3 (large mechanisms, more than 128 state variables): the block-per-reactor path with the fixed-pattern sparse LU
(csrc/rmsb_large.cuh, rmsb_sparse.h) against the oracle on a seeded synthetic mechanism (tests/synth.py; the reference ships
no mechanism of this size).  dydt / Jacobian / rate coefficients against the NumPy oracle (1e-12), trajectories against the
C++ CPU port of the same integrator with a DENSE pivoted LU (so the sparse no-pivot factorisation is checked end to end)."""
import numpy as np
import pytest

import cases
import rmsb200 as rb
import synth
from oracle import cpu_oracle as co
from oracle import rms_oracle as ro
from rmsb200 import _lib as L

pytestmark = pytest.mark.gpu
NS = 300


@pytest.fixture(scope="module")
def mech(tmp_path_factory):
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    path = str(tmp_path_factory.mktemp("synth") / f"synth{NS}.rms")
    synth.write_synthetic_mechanism(path, ns=NS, seed=7)
    d = rb.readinput(path)
    ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase", diffusionlimited=False)
    od = ro.readinput(path)
    oph = ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0], diffusionlimited=False)
    ic = {"T": 650.0, "V": 1.0, **synth.initial_moles(NS)}
    gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds=ic)
    odom = ro.ConstantTVDomain(oph, ic)
    return ph, oph, gdom, odom, y0, p


def test_large_pattern_analysis(mech):
    ph = mech[0]
    info = np.zeros(8, dtype=np.int64)
    L.check(L.lib().rmsb_large_info(ph.handle, L.ptr(info)))
    n, nnz, lvl_l, lvl_u, max_row, fma, core, scratch = (int(v) for v in info)
    assert n == NS and NS <= nnz <= NS * NS and 0 < core < NS
    assert lvl_l >= NS - core and lvl_u >= NS - core          # the dense core is a chain
    assert max_row <= NS and fma > 0 and scratch > 19 * NS


def test_large_dydt_jacobian_rates_parity(mech):
    ph, oph, gdom, odom, y0, p = mech
    rng = np.random.default_rng(1)
    ys = np.stack([odom.y0, np.abs(odom.y0 + 10.0 ** rng.uniform(-3, 1, NS)), np.abs(odom.y0 * rng.uniform(0.5, 1.5, NS) + 1e-6)])
    ens = rb.Domain(gdom.kind, ph, ys, np.repeat(gdom.cond, len(ys), axis=0))
    f = ens.dydt(ys)
    J = ens.jacobiany(ys)
    for k in range(len(ys)):
        assert cases.err_dydt(f[k], ro.dydtreactor(odom, ys[k], 0.0, None)) < 1e-12
        Jo = ro.jacobiany(odom, ys[k], 0.0, None)
        assert cases.err_jac(J[k], Jo) < 1e-12
        assert np.all(J[k][Jo == 0.0] == 0.0)
    np.testing.assert_allclose(gdom.kfs[0], odom.kfs, rtol=1e-12)
    np.testing.assert_allclose(gdom.krevs[0], odom.krevs, rtol=1e-12)
    np.testing.assert_allclose(p, odom.p, rtol=1e-12)
    # explicit parameters p = [G; k] give the same answer as the device's own tables
    q = np.array(p, dtype=float)
    q[NS + 3] *= 1.5
    f2 = ens.dydt(ys, q)
    assert cases.err_dydt(f2[0], ro.dydtreactor(odom, ys[0], 0.0, q)) < 1e-12


@pytest.mark.parametrize("tend,rtol,tol", [(1e-3, 1e-9, 1e-6), (1.0, 1e-8, 1e-5)])
def test_large_trajectory_vs_cpu_port(mech, tend, rtol, tol):
    ph, oph, gdom, odom, y0, p = mech
    tq = np.array([0.1 * tend, 0.5 * tend, tend])
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p).ode, abstol=1e-14, reltol=rtol, saveat=tq)
    assert sol.retcode == ["Success"]
    ref = co.CpuMech(oph).solve("ConstantTV", odom.y0[None], np.array([[650.0, 1.0]]), tend, rtol, 1e-14, tsave=tq, nthreads=1)
    assert ref["status"][0] == 0
    for k in range(len(tq)):
        yr = ref["ysave"][0, k]
        big = np.abs(yr) > 1e-8 * np.abs(yr).max()
        assert np.max(np.abs(sol.u[0, k] - yr)[big] / np.abs(yr)[big]) < tol
    # same algorithm, sparse vs dense linear algebra: the step counts agree
    assert abs(int(sol.stats["n_steps"][0]) - int(ref["stats"]["n_steps"][0])) <= 0.05 * ref["stats"]["n_steps"][0]


def test_large_ensemble_properties(mech):
    """Size-independent properties on an ensemble: all converge, 'mass' (the generator's conserved weights) is conserved,
    results do not depend on the position in the batch, constant species stay put."""
    ph, oph, gdom, odom, y0, p = mech
    n = 300
    T = np.random.default_rng(2).uniform(600.0, 700.0, n)
    y0s = np.repeat(gdom.y0, n, axis=0) * np.random.default_rng(3).uniform(0.8, 1.2, (n, 1))
    T[-1], y0s[-1] = T[0], y0s[0]
    ens = rb.Domain(gdom.kind, ph, y0s, np.stack([T, np.ones(n)], axis=1))
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 1.0)).ode, abstol=1e-12, reltol=1e-6)
    assert np.all(sol.status == 0)
    assert np.array_equal(sol.yend[0], sol.yend[-1])
    w = synth.species_weights(NS)
    np.testing.assert_allclose(sol.yend @ w, y0s @ w, rtol=1e-9)
    rev = rb.Domain(gdom.kind, ph, y0s[::-1].copy(), np.stack([T[::-1], np.ones(n)], axis=1))
    sol2 = rb.solve(rb.Reactor(rev, None, (0.0, 1.0)).ode, abstol=1e-12, reltol=1e-6)
    assert np.array_equal(sol2.yend[::-1], sol.yend)
    cs = rb.Domain(gdom.kind, ph, y0s[:4], np.stack([T[:4], np.ones(4)], axis=1), constantspecies=[f"S{NS - 1:04d}"])
    sol3 = rb.solve(rb.Reactor(cs, None, (0.0, 1e-2)).ode, abstol=1e-12, reltol=1e-6)
    assert np.all(sol3.status == 0) and np.array_equal(sol3.yend[:, NS - 1], y0s[:4, NS - 1])


def test_large_constant_tp(mech):
    """ConstantTP on the large path (y = [ns; V]: the species block bordered by a dense volume row and column)."""
    ph, oph, gdom, odom, y0, p = mech
    ic = {"T": 650.0, "P": 5.0e7, **synth.initial_moles(NS)}
    tdom, ty0, tp = rb.ConstantTPDomain(phase=ph, initialconds=ic)
    todom = ro.ConstantTPDomain(oph, ic)
    np.testing.assert_allclose(ty0, todom.y0, rtol=1e-14)
    np.testing.assert_allclose(tp, todom.p, rtol=1e-12)
    rng = np.random.default_rng(4)
    ys = np.stack([todom.y0, np.abs(todom.y0 * rng.uniform(0.5, 1.5, NS + 1) + 1e-6)])
    ens = rb.Domain(tdom.kind, ph, ys, np.repeat(tdom.cond, 2, axis=0))
    f, J = ens.dydt(ys), ens.jacobiany(ys)
    for k in range(2):
        assert cases.err_dydt(f[k], ro.dydtreactor(todom, ys[k], 0.0, None)) < 1e-12
        assert cases.err_jac(J[k], ro.jacobiany(todom, ys[k], 0.0, None)) < 1e-12
    tq = np.array([1e-4, 1e-3])
    sol = rb.solve(rb.Reactor(tdom, ty0, (0.0, 1e-3), p=tp).ode, abstol=1e-14, reltol=1e-9, saveat=tq)
    assert sol.retcode == ["Success"]
    ref = co.CpuMech(oph).solve("ConstantTP", todom.y0[None], np.array([[650.0, 5.0e7]]), 1e-3, 1e-9, 1e-14, tsave=tq, nthreads=1)
    assert ref["status"][0] == 0
    for k in range(len(tq)):
        yr = ref["ysave"][0, k]
        big = np.abs(yr) > 1e-8 * np.abs(yr[:NS]).max()
        assert np.max(np.abs(sol.u[0, k] - yr)[big] / np.abs(yr)[big]) < 1e-6
    assert sol.yend[0, NS] != ty0[NS]                                   # the volume moves


def test_large_path_limits_are_reported(mech):
    ph = mech[0]
    with pytest.raises(rb.RMSBError):
        _variable_k_domain_on_large_mechanism(ph, mech[2].y0)


def _variable_k_domain_on_large_mechanism(ph, y0):
    # a variable-k domain on more than 128 species is outside the large path; build the call through the C ABI directly
    y = np.concatenate([y0[0], [650.0, 1e5]])[None]
    out = np.empty_like(y)
    rc = L.lib().rmsb_dydt_batch(ph.handle, L.DOMAIN_CONSTANT_P, 1, L.ptr(np.ascontiguousarray(y)), L.ptr(np.array([[1e5]])), None, 0, None, 0, L.ptr(out))
    L.check(rc)


# ---------------------------------------------------------------------------------------------------------------------------------
# Config C3 AT ITS STATED SIZE (BASELINE.json configs[2], SURVEY 8d recipe): the reference's 35-species liquid octane mechanism
# replicated over 36 fuel families sharing the O2 / HO2 / OH pool -- 1015 species, 11 593 reactions, diffusion limited, ConstantTV.
# dydt / Jacobian / rate coefficients against the NumPy oracle (live and committed vectors), the trajectory against the committed
# SciPy-BDF golden (tests/golden/make_c3_golden.py), a 4096-condition sweep with zero non-OK statuses.
# ---------------------------------------------------------------------------------------------------------------------------------
K_FAM = 36


@pytest.fixture(scope="module")
def octane_families(tmp_path_factory):
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    path = str(tmp_path_factory.mktemp("c3") / "octane_families.rms")
    info = synth.write_octane_families(path, families=K_FAM)
    d = rb.readinput(path)
    ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase")
    od = ro.readinput(path)
    oph = ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0])
    ic = {"T": 450.0, "V": 1.0, **synth.octane_family_moles(K_FAM)}
    gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds=ic, constantspecies=["oxygen"])
    odom = ro.ConstantTVDomain(oph, ic, constantspecies=("oxygen",))
    gold = np.load(cases.ROOT / "tests" / "golden" / "c3_octane_families.npz")
    assert info["ns"] == int(gold["ns"]) == ph.Ns and info["nr"] == int(gold["nr"]) == ph.Nr
    return ph, oph, gdom, odom, y0, p, gold


def test_c3_stated_size_pattern(octane_families):
    ph = octane_families[0]
    assert ph.Ns >= 1000 and 10000 <= ph.Nr <= 12000
    info = np.zeros(8, dtype=np.int64)
    L.check(L.lib().rmsb_large_info(ph.handle, L.ptr(info)))
    n, nnz, lvl_l, lvl_u, max_row, fma, core, scratch = (int(v) for v in info)
    assert n == ph.Ns and 0.01 < nnz / n ** 2 < 0.04          # 1-4 % dense
    assert lvl_l < 80 and n - core < 64                        # families eliminate in parallel; the shared pool is the dense core


def test_c3_stated_size_dydt_jacobian_rates_parity(octane_families):
    from oracle import floors as fl
    ph, oph, gdom, odom, y0, p, gold = octane_families
    ys = gold["ys"]
    ens = rb.Domain(gdom.kind, ph, ys, np.repeat(gdom.cond, len(ys), axis=0), constantspecies=["oxygen"])
    ens.p, ens._p_default = gdom.p, gdom._p_default
    f, J = ens.dydt(ys, p), ens.jacobiany(ys, p)
    np.testing.assert_allclose(gdom.kfs[0], gold["kfs"], rtol=1e-12)
    np.testing.assert_allclose(gdom.krevs[0], gold["krevs"], rtol=1e-12)
    for k in range(len(ys)):
        fo, M = fl.dydt_terms(odom, ys[k], 0.0, odom.p)
        np.testing.assert_allclose(fo, gold["f"][k], rtol=1e-12, atol=1e-30)       # the committed vectors are the live oracle's
        rep = fl.report(f[k], fo, fl.EPS * M, 1e-12, slack=4.0)
        assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, rep
        Jo = ro.jacobiany(odom, ys[k], 0.0, odom.p)
        assert cases.err_jac(J[k], Jo) < 1e-12
        nz = Jo != 0.0
        # species-block entries are short sums of same-signed or few terms: plain relative error on every non-zero entry
        assert np.max(np.abs(J[k][nz] - Jo[nz]) / np.abs(Jo[nz])) < 1e-9
        assert np.all(J[k][~nz] == 0.0)


def test_c3_stated_size_trajectory_vs_scipy_golden(octane_families):
    ph, oph, gdom, odom, y0, p, gold = octane_families
    tq, yq = gold["tq"], gold["yq"]
    # The golden is SciPy's BDF at rtol 1e-9 over the NumPy oracle (11 CPU-minutes; tighter settings do not finish in hours), so it
    # carries its own global error.  Measured (scripts/gpu_c3_golden_check.py): the GPU trajectory converges on itself -- rtol 1e-9 /
    # 1e-10 / 1e-11 sit 4.9e-7 / 9e-8 / 6e-9 from the rtol 1e-12 run -- while its distance to the golden saturates at 1.15e-6 from
    # rtol 1e-10 on: that distance is the golden's error.  Bars: 1e-6 between the GPU at rtol 1e-9 and at rtol 1e-11, 2e-6 to SciPy.
    def dist(a, b):
        out = 0.0
        for k in range(len(tq)):
            big = np.abs(b[k]) > 1e-9 * np.abs(b[k]).max()
            out = max(out, float(np.max(np.abs(a[k] - b[k])[big] / np.abs(b[k])[big])))
        return out

    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, float(tq[-1])), p=p).ode, abstol=1e-18, reltol=1e-9, saveat=tq)
    sol11 = rb.solve(rb.Reactor(gdom, y0, (0.0, float(tq[-1])), p=p).ode, abstol=1e-18, reltol=1e-11, saveat=tq)
    assert sol.retcode == ["Success"] and sol11.retcode == ["Success"]
    assert dist(sol.u[0], sol11.u[0]) < 1e-6
    assert dist(sol11.u[0], yq) < 2e-6
    i_oct = ph.names.index("F00_octane")
    assert sol.u[0, -1, i_oct] < 0.5 * y0[i_oct]                               # the fuel is mostly converted by t_end: not a trivial trajectory
    # the reference's loose setting lands on the same solution
    sol6 = rb.solve(rb.Reactor(gdom, y0, (0.0, float(tq[-1])), p=p).ode, abstol=1e-16, reltol=1e-6, saveat=tq[-1:])
    big = np.abs(yq[-1]) > 1e-6 * np.abs(yq[-1]).max()
    assert np.max(np.abs(sol6.u[0, 0] - yq[-1])[big] / np.abs(yq[-1])[big]) < 2e-3


def test_c3_stated_size_4k_sweep_all_converge(octane_families):
    """BASELINE configs[2]: the 4k-condition sweep.  Zero non-OK statuses, O2 held, every family's carbon skeleton count conserved
    (each family reaction maps family species to family species one to one), batch position does not matter."""
    ph, oph, gdom, odom, y0, p, gold = octane_families
    n = 4096
    rng = np.random.default_rng(cases.SEED)
    T = rng.uniform(400.0, 500.0, n)
    y0s = np.repeat(gdom.y0, n, axis=0) * rng.uniform(0.8, 1.2, (n, 1))
    io2 = ph.names.index("oxygen")
    y0s[:, io2] = 4.953 * rng.uniform(0.5, 2.0, n)
    T[-1], y0s[-1] = T[0], y0s[0]
    ens = rb.Domain(gdom.kind, ph, y0s, np.stack([T, np.ones(n)], axis=1), constantspecies=["oxygen"])
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 140000.01)).ode, abstol=1e-16, reltol=1e-6)
    assert np.array_equal(np.bincount(sol.status, minlength=4), [n, 0, 0, 0])
    assert np.array_equal(sol.yend[:, io2], y0s[:, io2])
    assert np.array_equal(sol.yend[0], sol.yend[-1])
    fam = np.array([int(nm[1:3]) if nm.startswith("F") and nm[3] == "_" else -1 for nm in ph.names])
    for k in (0, 17, K_FAM - 1):
        np.testing.assert_allclose(sol.yend[:, fam == k].sum(axis=1), y0s[:, fam == k].sum(axis=1), rtol=1e-9)


# ---------------------------------------------------------------------------------------------------------------------------------
# Large mechanisms on the VARIABLE-k gas domains (ConstantV / ConstantP: thermo and rate coefficients inside every RHS call, two
# dense border rows / columns for T and P | V in the sparse factorisation) and with third-body / falloff kinetics (efficiency
# entries in the fixed pattern; ConstantTP re-evaluates those reactions on every call) -- src/Domain.jl:1064-1113,1141-1217,
# 2151-2167, src/Reactor.jl:342-382.  The mechanism is the octane-family mechanism read as an IdealGas with its association
# reactions turned into Troe / Lindemann / ThirdBody forms (tests/synth.py, falloff=True).
# ---------------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module", params=[12, 36], ids=["343 species", "1015 species"])
def gas_families(request, tmp_path_factory):
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    K = request.param
    path = str(tmp_path_factory.mktemp("c3gas") / f"octane_gas_{K}.rms")
    synth.write_octane_families(path, families=K, falloff=True)
    d = rb.readinput(path)
    ph = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
    od = ro.readinput(path)
    oph = ro.IdealGas(od["phase"]["Species"], od["phase"]["Reactions"])
    fuel = {k: v / 6154.0 for k, v in synth.octane_family_moles(K).items() if k != "oxygen"}
    ic = {"T": 1000.0, "P": 20e5, **fuel, "oxygen": 6.25, "N2": 23.5}
    assert ph.Ns > 128 and len(ph.efficiencyinds) == len(oph.efficiencyinds) > 0
    return K, ph, oph, ic


@pytest.mark.parametrize("kind", ["ConstantV", "ConstantP", "ConstantTP"])
def test_large_gas_domains_parity(gas_families, kind):
    from oracle import floors as fl
    K, ph, oph, ic = gas_families
    gdom, y0, p = getattr(rb, kind + "Domain")(phase=ph, initialconds=ic, constantspecies=["N2"])
    odom = getattr(ro, kind + "Domain")(oph, ic, constantspecies=("N2",))
    Ns, n = ph.Ns, len(odom.y0)
    rng = np.random.default_rng(K)
    ys = np.stack([odom.y0, np.abs(odom.y0 * rng.uniform(0.5, 1.5, n) + np.r_[1e-6 * np.ones(Ns), np.zeros(n - Ns)])])
    ens = rb.Domain(gdom.kind, ph, ys, np.repeat(gdom.cond, len(ys), axis=0), constantspecies=["N2"])
    ens.p, ens._p_default = gdom.p, gdom._p_default
    f, J = ens.dydt(ys, p), ens.jacobiany(ys, p)
    kf, kr = ens.rates(ys, p)
    for k in range(len(ys)):
        fo, M = fl.dydt_terms(odom, ys[k], 0.0, odom.p)
        rep = fl.report(f[k], fo, fl.EPS * M, 1e-12, slack=4.0)
        assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 1.0, rep
        th = odom.calcthermo(ys[k], 0.0, odom.p)
        np.testing.assert_allclose(kf[k], th.kfs, rtol=1e-12)
        np.testing.assert_allclose(kr[k], th.krevs, rtol=1e-12, atol=1e-300)
        if k == 1 or K == 12:
            Jo, Jf = fl.jac_floor(lambda yy, pp: ro.jacobiany(odom, yy, 0.0, pp), odom, ys[k], odom.p, nsample=8)
            rep = fl.report(J[k], Jo, Jf, 1e-12)
            # the conditioning floor is itself an estimate from eight random perturbations (a factor ~2 either way on 10^6 entries)
            assert rep["max_rel_above"] <= 1e-12 and rep["worst_excess"] <= 2.0, [rep[q] for q in ("max_rel_above", "worst_excess", "worst_at", "worst_x", "worst_ref", "worst_floor")]
            assert np.all(J[k][ph.names.index("N2")] == 0.0)                    # constant-species row


def test_large_gas_trajectory_and_sweep(gas_families):
    """ConstantV trajectory against the C++ CPU port (same integrator, DENSE pivoted LU) and a sweep: all converge, the inert and the
    held species stay put, the batch position does not matter."""
    K, ph, oph, ic = gas_families
    gdom, y0, p = rb.ConstantVDomain(phase=ph, initialconds=ic)
    odom = ro.ConstantVDomain(oph, ic)
    Ns = ph.Ns
    tend = 2e-3
    tq = np.array([0.25 * tend, tend])
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, tend), p=p).ode, abstol=1e-16, reltol=1e-8, saveat=tq)
    assert sol.retcode == ["Success"]
    if K == 12:
        ref = co.CpuMech(oph).solve("ConstantV", odom.y0[None], odom.cond[None], tend, 1e-8, 1e-16, tsave=tq, nthreads=1)
        assert ref["status"][0] == 0
        for k in range(len(tq)):
            yr = ref["ysave"][0, k]
            big = np.abs(yr) > 1e-9 * np.abs(yr[:Ns]).max()
            assert np.max(np.abs(sol.u[0, k] - yr)[big] / np.abs(yr)[big]) < 2e-6
    n = 148
    rng = np.random.default_rng(5)
    T0 = rng.uniform(900.0, 1200.0, n); P0 = 10 ** rng.uniform(5.7, 6.7, n)
    T0[-1], P0[-1] = T0[0], P0[0]
    x = gdom.y0[0, :Ns] / gdom.y0[0, :Ns].sum()
    ens = rb.ConstantVEnsemble(ph, T0, P0, np.repeat(x[None], n, axis=0))
    s = rb.solve(rb.Reactor(ens, None, (0.0, 1e-2)).ode, abstol=1e-16, reltol=1e-6)
    assert np.array_equal(np.bincount(s.status, minlength=4), [n, 0, 0, 0])
    assert np.array_equal(s.yend[0], s.yend[-1])
    for inert in ("N2", "Ar"):
        i = ph.names.index(inert)
        np.testing.assert_allclose(s.yend[:, i], ens.y0[:, i], rtol=1e-12, atol=1e-300)
    # constant pressure: adiabatic, volume is a state
    pdom, py0, pp = rb.ConstantPDomain(phase=ph, initialconds=ic)
    sp = rb.solve(rb.Reactor(pdom, py0, (0.0, tend), p=pp).ode, abstol=1e-16, reltol=1e-8)
    assert sp.retcode == ["Success"] and sp.yend[0, Ns] > 1000.0 and sp.yend[0, Ns + 1] != py0[Ns + 1]
