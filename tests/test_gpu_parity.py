"""Parity tests proper: the CUDA path (through the C ABI of librmsb200.so) against the oracle on the same
seeded inputs, against the committed golden fixtures, and -- at ensemble sizes -- through size-independent
properties.  Tolerances (BASELINE.json north_star): dydt / Jacobian 1e-12 relative in FP64, species
trajectories 1e-6 relative at output times, ignition delay 0.1 %.
"""
import ctypes as C

import numpy as np
import pytest

import cases
import rmsb200 as rb
from oracle import cpu_oracle as co
from oracle import rms_oracle as ro
from rmsb200 import _lib as L

pytestmark = pytest.mark.gpu
G = np.load(cases.ROOT / "tests" / "golden" / "parity_vectors.npz")
TOL = 1e-12


def ensemble_of(gdom, ys):
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.constantspeciesinds = gdom.constantspeciesinds
    ens.p, ens._p_default = gdom.p, gdom._p_default
    return ens


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"


@pytest.mark.parametrize("case", list(cases.CASES))
def test_dydt_and_jacobian_parity(case):
    """CUDA dydt / jacobiany vs the committed oracle vectors AND vs the oracle evaluated live."""
    odom = cases.oracle_domain(case)
    gdom, y0, p = cases.product_domain(case)
    ys = G[f"{case}/y"]
    ens = ensemble_of(gdom, ys)
    f = ens.dydt(ys, p)
    J = ens.jacobiany(ys, p)
    po = odom.p if odom.kind in ("ConstantV", "ConstantTP") else None
    for k in range(len(ys)):
        assert cases.err_dydt(f[k], G[f"{case}/f"][k]) < TOL
        assert cases.err_jac(J[k], G[f"{case}/J"][k]) < TOL
    fo = ro.dydtreactor(odom, ys[0], 0.0, po)
    assert cases.err_dydt(f[0], fo) < TOL
    # plain elementwise relative error on every entry that is not cancellation noise
    Jo = G[f"{case}/J"][0]
    sig = np.abs(Jo) > 1e-6 * np.max(np.abs(Jo), axis=0, keepdims=True)
    assert np.max(np.abs(J[0] - Jo)[sig] / np.abs(Jo)[sig]) < 1e-9
    # structural zeros: entries the oracle has exactly zero are zero here too, up to cancellation residue of
    # +d -d pairs summed in a different order (bounded by 1e-15 of the column scale)
    col = np.max(np.abs(Jo), axis=0, keepdims=True) * np.ones_like(Jo)
    assert np.all(np.abs(J[0][Jo == 0.0]) <= 1e-15 * col[Jo == 0.0])


@pytest.mark.parametrize("case", ["superminimal_V", "minimal_V", "ethane_V", "propane_V"])
def test_dydt_with_parameters(case):
    """p = [dG; f] perturbations of the ConstantV domain (src/Domain.jl:1100,1112)."""
    gdom, y0, p = cases.product_domain(case)
    ys, pp = G[f"{case}/y"], G[f"{case}/p"]
    ens = ensemble_of(gdom, ys)
    f = ens.dydt(ys, pp)
    for k in range(len(ys)):
        assert cases.err_dydt(f[k], G[f"{case}/f_p"][k]) < TOL
    # per-reactor parameter rows give the same answer as the shared vector
    f2 = ens.dydt(ys, np.repeat(pp[None, :], len(ys), axis=0))
    assert np.array_equal(f, f2)


def test_constant_k_domain_parameters_match_oracle():
    """p = [Gs; kfs] returned by the constant-k domain constructors (src/Domain.jl:88-93,549-551,700-702)."""
    for case in ("superminimal_TP", "ethane_TP", "surface_TAPhi"):
        odom = cases.oracle_domain(case)
        gdom, y0, p = cases.product_domain(case)
        np.testing.assert_allclose(p, odom.p, rtol=1e-12)
        np.testing.assert_allclose(gdom.krevs[0], odom.krevs, rtol=1e-12)
    odom = cases.oracle_domain("liquid_TV")           # diffusion-limited liquid (src/PhaseState.jl:184-223)
    gdom, y0, p = cases.product_domain("liquid_TV")
    np.testing.assert_allclose(gdom.kfs[0], odom.kfs, rtol=1e-12)
    np.testing.assert_allclose(gdom.krevs[0], odom.krevs, rtol=1e-12)
    np.testing.assert_allclose(gdom.Gs[0], odom.Gs, rtol=1e-13)


def test_c2_ignition_delay_golden():
    """C2: IDT = 0.038384723436228063 s with the reference's own grid definition (src/TestReactors.jl:301-310)."""
    gdom, y0, p = cases.product_domain("superminimal_V")
    ts = np.exp(np.linspace(np.log(1e-15), np.log(0.1), 10000))
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.101), p=p), abstol=1e-20, reltol=1e-12, saveat=ts)   # the reference's tolerances
    assert sol.retcode == ["Success"]
    assert rb.idt_grid(sol) == pytest.approx(0.038384723436228063, rel=1e-5)
    assert float(G["superminimal_V/idt_grid"]) == rb.idt_grid(sol)
    sol10 = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.101), p=p), abstol=1e-20, reltol=1e-10, saveat=ts)
    assert rb.idt_grid(sol10) == rb.idt_grid(sol)
    # continuous definition agrees with the grid one to within a grid cell (0.32 %)
    assert sol.idt[0] == pytest.approx(rb.idt_grid(sol), rel=4e-3)


def test_c1_constant_tp_golden():
    """C1: x(H2), x(O2), x(H2O) at t = 20.44002454 and the ROP of H2 by reaction 1 (src/TestReactors.jl:193-213)."""
    gdom, y0, p = cases.product_domain("superminimal_TP")
    t = 20.44002454
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 150.11094), p=p), abstol=1e-20, reltol=1e-10, saveat=[1.0, t, 100.0, 150.11094])
    sim = rb.Simulation(sol, gdom)
    assert sim.molefractions("H2", t) == pytest.approx(0.412883111, rel=1e-4)
    assert sim.molefractions("O2", t) == pytest.approx(0.200419093, rel=1e-4)
    assert sim.molefractions("H2O", t) == pytest.approx(0.386618602, rel=1e-4)
    assert sim.rops("H2", t)[0] == pytest.approx(-5.304207075533127e-13, rel=1e-4)
    yq = G["superminimal_TP/yq"]
    big = np.abs(yq) > 1e-12 * np.max(np.abs(yq[:, :12]))
    assert np.max(np.abs(sol.u[0] - yq)[big] / np.abs(yq)[big]) < 1e-6


def test_trajectory_parity_tight_tolerance():
    """Species trajectories within 1e-6 relative at output times vs the oracle (SciPy BDF, rtol 1e-10)."""
    gdom, y0, p = cases.product_domain("superminimal_V")
    tq = G["superminimal_V/tq"]
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.101), p=p), abstol=1e-20, reltol=1e-10, saveat=tq)
    yq = G["superminimal_V/yq"]
    big = np.abs(yq) > 1e-12 * np.max(np.abs(yq[:, :12]))
    assert np.max(np.abs(sol.u[0] - yq)[big] / np.abs(yq)[big]) < 1e-6


def sweep(ig, n, seed=cases.SEED):
    """SURVEY 8(d) C2 condition sweep."""
    rng = np.random.default_rng(seed)
    T0, P0, phi = rng.uniform(900, 1500, n), 10 ** rng.uniform(5, 6.7, n), rng.uniform(0.5, 2.0, n)
    ns = np.zeros((n, ig.Ns))
    ns[:, ig.spcdict["H2"] - 1] = 2 * phi
    ns[:, ig.spcdict["O2"] - 1] = 1.0
    ns[:, ig.spcdict["N2"] - 1] = 3.76 * rng.integers(0, 2, n)
    ns /= ns.sum(axis=1, keepdims=True)
    return T0, P0, ns


def test_ensemble_vs_cpu_oracle_idt():
    """64 conditions of the C2 sweep: CUDA ensemble vs the C++ CPU oracle, both at rtol 1e-10 / atol 1e-20:
    ignition delay within 0.1 %, final state within 1e-6 on every significant species."""
    gdom, _, _ = cases.product_domain("superminimal_V")
    ig = gdom.phase
    T0, P0, ns = sweep(ig, 64)
    ens = rb.ConstantVEnsemble(ig, T0, P0, ns)
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-20, reltol=1e-10)
    odom = cases.oracle_domain("superminimal_V")
    ref = co.CpuMech(odom.phase).solve("ConstantV", ens.y0, ens.cond, 0.1, 1e-10, 1e-20)
    assert np.all(sol.status == 0) and np.all(ref["status"] == 0)
    ign = (sol.yend[:, ig.Ns] - ens.y0[:, ig.Ns]) > 400.0
    assert ign.sum() > 10
    np.testing.assert_allclose(sol.idt[ign], ref["idt"][ign], rtol=1e-3)
    scale = np.max(np.abs(ref["yend"][:, :ig.Ns]), axis=1, keepdims=True)
    big = np.abs(ref["yend"]) > 1e-9 * scale
    assert np.max(np.abs(sol.yend - ref["yend"])[big] / np.abs(ref["yend"])[big]) < 1e-6


def test_c2_subsample_vs_scipy():
    """The same 64 conditions against SciPy BDF over the NumPy oracle (tests/golden/make_c2_scipy_subsample.py, rtol 1e-10 /
    atol 1e-20): an integrator and a right-hand side that share nothing with the CUDA path or the C++ port.  Ignition delay within
    0.1 % (north_star), final state within 1e-6 on every significant species; and at the throughput tolerances the delay still
    within 0.1 %."""
    gold = np.load(cases.ROOT / "tests" / "golden" / "c2_scipy_subsample.npz")
    gdom, _, _ = cases.product_domain("superminimal_V")
    ig = gdom.phase
    T0, P0, ns = sweep(ig, 64)
    ens = rb.ConstantVEnsemble(ig, T0, P0, ns)
    np.testing.assert_allclose(ens.y0, gold["y0"], rtol=1e-13)
    np.testing.assert_allclose(ens.cond, gold["cond"].reshape(ens.cond.shape), rtol=1e-13)
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-20, reltol=1e-10)
    assert np.all(sol.status == 0)
    yr = gold["yend"]
    ign = (yr[:, ig.Ns] - ens.y0[:, ig.Ns]) > 400.0
    assert ign.sum() > 10
    np.testing.assert_allclose(sol.idt[ign], gold["idt"][ign], rtol=1e-3)
    scale = np.max(np.abs(yr[:, :ig.Ns]), axis=1, keepdims=True)
    big = np.abs(yr) > 1e-9 * scale
    assert np.max(np.abs(sol.yend - yr)[big] / np.abs(yr)[big]) < 1e-6
    loose = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-16, reltol=1e-6)
    np.testing.assert_allclose(loose.idt[ign], gold["idt"][ign], rtol=1e-3)


def test_ensemble_loose_tolerance_vs_tight_cpu_oracle():
    """2 000 conditions at the THROUGHPUT tolerances (rtol 1e-6, atol 1e-16) against the C++ CPU oracle at rtol 1e-10:
    ignition delay within 0.1 % (north_star), final temperature within 1e-4."""
    gdom, _, _ = cases.product_domain("superminimal_V")
    ig = gdom.phase
    T0, P0, ns = sweep(ig, 2000)
    ens = rb.ConstantVEnsemble(ig, T0, P0, ns)
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-16, reltol=1e-6)
    odom = cases.oracle_domain("superminimal_V")
    ref = co.CpuMech(odom.phase).solve("ConstantV", ens.y0, ens.cond, 0.1, 1e-10, 1e-20)
    assert np.all(sol.status == 0)
    ign = (ref["yend"][:, ig.Ns] - ens.y0[:, ig.Ns]) > 400.0
    assert ign.sum() > 1000
    np.testing.assert_allclose(sol.idt[ign], ref["idt"][ign], rtol=1e-3)
    np.testing.assert_allclose(sol.yend[:, ig.Ns], ref["yend"][:, ig.Ns], rtol=1e-4)
    # the CPU port of the same algorithm at the same loose tolerance is no more accurate than the CUDA path
    ref6 = co.CpuMech(odom.phase).solve("ConstantV", ens.y0, ens.cond, 0.1, 1e-6, 1e-16)
    e_gpu = np.max(np.abs(sol.yend[:, ig.Ns] / ref["yend"][:, ig.Ns] - 1))
    e_cpu = np.max(np.abs(ref6["yend"][:, ig.Ns] / ref["yend"][:, ig.Ns] - 1))
    assert e_cpu < 5e-5 and e_gpu < 5e-5      # both are tolerance-limited (rtol 1e-6 over ~400 steps through ignition)
    # the step counts track the CPU implementation of the same algorithm, with the production defaults (flags 0) and with
    # the scipy-compatible Newton test / step logic (flags 3)
    assert abs(sol.stats["n_steps"].mean() / ref6["stats"]["n_steps"].mean() - 1.0) < 0.05
    assert abs(sol.stats["n_rhs"].mean() / ref6["stats"]["n_rhs"].mean() - 1.0) < 0.05
    ref3 = co.CpuMech(odom.phase).solve("ConstantV", ens.y0, ens.cond, 0.1, 1e-6, 1e-16, flags=3)
    sol3 = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), rb.RMSB200(flags=3), abstol=1e-16, reltol=1e-6)
    assert abs(sol3.stats["n_steps"].mean() / ref3["stats"]["n_steps"].mean() - 1.0) < 0.05
    assert abs(sol3.stats["n_rhs"].mean() / ref3["stats"]["n_rhs"].mean() - 1.0) < 0.05
    # the production defaults (looser Newton tolerance, step hysteresis) need fewer RHS evaluations
    assert sol.stats["n_rhs"].mean() < 0.8 * sol3.stats["n_rhs"].mean()


def test_ensemble_properties_at_scale():
    """Size-independent properties on a 20 000-reactor ensemble at the throughput tolerances (rtol 1e-6, atol 1e-16):
    element conservation, batch-order independence, determinism, agreement of duplicates."""
    gdom, _, _ = cases.product_domain("superminimal_V")
    ig = gdom.phase
    n = 20000
    T0, P0, ns = sweep(ig, n)
    T0[-1], P0[-1], ns[-1] = T0[0], P0[0], ns[0]                      # a duplicate condition
    ens = rb.ConstantVEnsemble(ig, T0, P0, ns)
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-16, reltol=1e-6)
    assert np.all(sol.status == 0)
    assert np.array_equal(sol.yend[0], sol.yend[-1]) and sol.idt[0] == sol.idt[-1]
    # H and O atoms are conserved by every reaction: check the element balance of the final state
    names = ig.names
    nH = np.array([{"H2": 2, "H": 1, "OH": 1, "HO2": 1, "H2O2": 2, "H2O": 2}.get(s, 0) for s in names], dtype=float)
    nO = np.array([{"O2": 2, "O(T)": 1, "OH": 1, "HO2": 2, "H2O2": 2, "H2O": 1}.get(s, 0) for s in names], dtype=float)
    for nel in (nH, nO):       # linear invariants are preserved by BDF + exact linear solves to rounding
        a0, a1 = ens.y0[:, :ig.Ns] @ nel, sol.yend[:, :ig.Ns] @ nel
        np.testing.assert_allclose(a1, a0, rtol=1e-12)
    # reversed batch order gives bitwise the same per-reactor results (no cross-reactor coupling, deterministic kernels)
    rev = rb.ConstantVEnsemble(ig, T0[::-1].copy(), P0[::-1].copy(), ns[::-1].copy())
    sol2 = rb.solve(rb.Reactor(rev, None, (0.0, 0.1)), abstol=1e-16, reltol=1e-6)
    assert np.array_equal(sol2.yend[::-1], sol.yend) and np.array_equal(sol2.idt[::-1], sol.idt, equal_nan=True)
    # ... and so does keeping the caller's order in the work queue (flags bit 6) instead of the library's hottest-first order
    sol3 = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), rb.RMSB200(flags=64), abstol=1e-16, reltol=1e-6)
    assert np.array_equal(sol3.yend, sol.yend) and np.array_equal(sol3.idt, sol.idt, equal_nan=True)
    # the ideal-gas law is an invariant of the ConstantV equations (dP/dt is derived from it, Domain.jl:1903):
    # P*V/(R*T) must track sum(ns) to integration accuracy (the CPU port drifts by 6e-7 at these tolerances).
    Ni = sol.yend[:, ig.Ns + 1] * ens.cond[:, 0] / (8.314 * sol.yend[:, ig.Ns])
    np.testing.assert_allclose(Ni, sol.yend[:, :ig.Ns].sum(axis=1), rtol=1e-5)


def test_edge_cases():
    gdom, y0, p = cases.product_domain("superminimal_V")
    lib = L.lib()
    # empty batch is a no-op
    assert lib.rmsb_dydt_batch(gdom.phase.handle, gdom.kind, 0, None, None, None, 0, None, 0, None) == 0
    # bad domain / null buffers are rejected with a message, not a crash
    assert lib.rmsb_dydt_batch(gdom.phase.handle, 77, 1, L.ptr(gdom.y0), L.ptr(gdom.cond), None, 0, None, 0, L.ptr(gdom.y0.copy())) == -1
    assert b"domain" in lib.rmsb_last_error()
    # inert mixture: nothing happens, integration reaches t_end in a handful of steps
    ig = gdom.phase
    ns = np.zeros((1, ig.Ns)); ns[0, ig.spcdict["N2"] - 1] = 1.0
    ens = rb.ConstantVEnsemble(ig, [1000.0], [1e5], ns)
    sol = rb.solve(rb.Reactor(ens, None, (0.0, 1.0)), abstol=1e-16, reltol=1e-6, saveat=[0.0, 0.5, 1.0])
    assert sol.status[0] == 0 and np.allclose(sol.u[0, :, :], ens.y0[0], rtol=1e-12)
    # max_steps exhaustion is reported per reactor, never a global abort
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 0.101), p=p), rb.RMSB200(max_steps=10), abstol=1e-20, reltol=1e-10)
    assert sol.status[0] == L.SOLVE_MAX_STEPS and sol.stats["n_steps"][0] == 10
    # broken mechanism (src/TestReactors.jl:630-647 fault-injection fixture): finishes or reports, never hangs
    d = rb.readinput(cases.fixture("superminimal_broken"))
    igb = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"])
    domb, y0b, pb = rb.ConstantVDomain(phase=igb, initialconds={"T": 1000.0, "P": 10e5, "H2": 0.67, "O2": 0.33})
    solb = rb.solve(rb.Reactor(domb, y0b, (0.0, 0.101), p=pb), rb.RMSB200(max_steps=20000), abstol=1e-16, reltol=1e-6)
    assert solb.status[0] in (0, 1, 2, 3)


def test_liquid_octane_golden():
    """C3 (shipped fixture): x(octane) = 0.461599061 at t = 32977.61568 s, rtol 3e-2 (src/TestReactors.jl:18-27),
    and parity of the trajectory with the C++ CPU oracle."""
    gdom, y0, p = cases.product_domain("liquid_TV")
    t = 32977.61568
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 140000.01), p=p), abstol=1e-20, reltol=1e-8, saveat=[t, 140000.01])
    assert sol.status[0] == 0
    y = sol.u[0, 0]
    x = y[gdom.phase.spcdict["octane"] - 1] / y.sum()
    assert x == pytest.approx(0.461599061, rel=3e-2)
    odom = cases.oracle_domain("liquid_TV")
    ref = co.CpuMech(odom.phase).solve("ConstantTV", odom.y0, odom.cond, 140000.01, 1e-8, 1e-20,
                                       constspc=odom.constantspeciesinds, tsave=[t, 140000.01])
    big = np.abs(ref["ysave"][0]) > 1e-6 * np.max(np.abs(ref["ysave"][0]))
    assert np.max(np.abs(sol.u[0] - ref["ysave"][0])[big] / np.abs(ref["ysave"][0])[big]) < 1e-5


def test_device_resident_entry_points_match_host_ones():
    """The *_dev entry points (device pointers, caller's stream) give bitwise the same results as the host ones."""
    import torch
    gdom, y0, p = cases.product_domain("superminimal_V")
    ys = G["superminimal_V/y"]
    ens = ensemble_of(gdom, ys)
    f_host = ens.dydt(ys)
    dev = torch.device("cuda:0")
    ty, tc = torch.from_numpy(ys).to(dev), torch.from_numpy(ens.cond).to(dev)
    tf = torch.empty_like(ty)
    st = torch.cuda.current_stream().cuda_stream
    L.check(L.lib().rmsb_dydt_batch_dev(gdom.phase.handle, gdom.kind, len(ys), ty.data_ptr(), tc.data_ptr(), None, 0, None, 0,
                                        tf.data_ptr(), C.c_void_p(st)))
    torch.cuda.synchronize()
    assert np.array_equal(tf.cpu().numpy(), f_host)


def test_three_body_diffusion_limits(tmp_path):
    """Generalised Smoluchowski rate for three reacting bodies (getDiffusiveRate, src/PhaseState.jl:92-101) inside getkfkrev's
    diffusion limits: the liquid octane fixture with a 3 -> 1, a 1 -> 3 and a 3 -> 3 reaction added; kf / krev, dydt and the Jacobian
    against the oracle."""
    import yaml
    with open(cases.fixture("liquid_phase")) as fh:
        doc = yaml.load(fh, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))
    extra = [(["octane", "oxygen", "OH(D)"], ["CCCCCCC(C)OO"], 3.0e2), (["CCCCC(CCC)OO"], ["octane", "oxygen", "[O]O"], 5.0e4),
             (["octane", "[O]O", "OH(D)"], ["CC[CH]CCCCC", "oxygen", "CCCCCCC(C)OO"], 8.0e3)]
    for re_, pr_, A in extra:
        doc["Reactions"].append({"kinetics": {"A": A, "Ea": 2.0e4, "n": 0.3, "type": "Arrhenius"}, "products": pr_, "radicalchange": 0,
                                 "reactants": re_, "type": "ElementaryReaction"})
    path = str(tmp_path / "liquid3.rms")
    with open(path, "w") as fh:
        yaml.safe_dump(doc, fh)
    ic = {"T": 450.0, "V": 1.0, "octane": 6154.0, "oxygen": 4.953}
    d = rb.readinput(path)
    ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase")
    od = ro.readinput(path)
    oph = ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0])
    gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds=ic, constantspecies=["oxygen"])
    odom = ro.ConstantTVDomain(oph, ic, constantspecies=("oxygen",))
    assert ph.Nr == 383
    three = [j for j, r in enumerate(oph.reactions) if len(r.reactantinds) == 3 or len(r.productinds) == 3]
    assert len(three) == 3
    np.testing.assert_allclose(gdom.kfs[0], odom.kfs, rtol=1e-12)
    np.testing.assert_allclose(gdom.krevs[0], odom.krevs, rtol=1e-12)
    kf_nolimit = ro.ConstantTVDomain(ro.IdealDiluteSolution(od["phase"]["Species"], od["phase"]["Reactions"], od["Solvents"][0], diffusionlimited=False),
                                     ic, constantspecies=("oxygen",)).kfs
    assert np.any(np.abs(odom.kfs[three] - kf_nolimit[three]) > 1e-6 * kf_nolimit[three])     # the three-body limit bites
    ys = cases.states("liquid_TV", odom.y0, oph.Ns, "ConstantTV")[:2]
    ens = ensemble_of(gdom, ys)
    f, J = ens.dydt(ys, p), ens.jacobiany(ys, p)
    for k in range(len(ys)):
        assert cases.err_dydt(f[k], ro.dydtreactor(odom, ys[k], 0.0, None)) < TOL
        assert cases.err_jac(J[k], ro.jacobiany(odom, ys[k], 0.0, None)) < TOL
