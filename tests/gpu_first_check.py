"""Exploratory GPU check (not a pytest file): prints CUDA-vs-oracle differences for every domain.
Run on the GPU box:  python tests/gpu_first_check.py
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "reactionmechanismsimulator.jl_b200"))

import rmsb200 as rb  # noqa: E402
from oracle import rms_oracle as ro  # noqa: E402
from oracle import integrate as oi  # noqa: E402

FX = ROOT / "tests" / "fixtures"
rng = np.random.default_rng(20261019)


def relerr(a, b):
    scale = np.maximum(np.abs(b), 1e-300)
    return np.max(np.abs(a - b) / scale)


def mixed(a, b, axis_scale):
    """max |a-b| / (|b| + 1e-3*column scale)"""
    return np.max(np.abs(a - b) / (np.abs(b) + 1e-3 * axis_scale + 1e-300))


def check_domain(tag, odom, gdom, p_o, p_g, nstate=8):
    Ns = odom.phase.Ns
    ys = []
    for k in range(nstate):
        y = odom.y0.copy()
        y[:Ns] = np.abs(y[:Ns] + y[:Ns].sum() * 10.0 ** rng.uniform(-8, -2, Ns) * rng.uniform(0.1, 1, Ns))
        if odom.kind == "ConstantV":
            y[Ns] *= rng.uniform(0.9, 1.6)
            y[Ns + 1] *= rng.uniform(0.8, 1.5)
        ys.append(y)
    ys = np.array(ys)
    gd = type(gdom)(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, nstate, axis=0), ())
    gd.constantspeciesinds = gdom.constantspeciesinds
    gd._p_default = gdom._p_default
    t0 = time.time()
    f_g = gd.dydt(ys, p_g)
    J_g = gd.jacobiany(ys, p_g)
    tg = time.time() - t0
    worst_f = worst_J = worst_Jm = 0.0
    for k in range(nstate):
        f_o = ro.dydtreactor(odom, ys[k], 0.0, p_o)
        J_o = ro.jacobiany(odom, ys[k], 0.0, p_o)
        fs = np.max(np.abs(f_o))
        worst_f = max(worst_f, np.max(np.abs(f_g[k] - f_o) / (np.abs(f_o) + 1e-6 * fs + 1e-300)))
        colscale = np.max(np.abs(J_o), axis=0, keepdims=True)
        worst_Jm = max(worst_Jm, mixed(J_g[k], J_o, colscale))
        nz = np.abs(J_o) > 1e-8 * colscale
        worst_J = max(worst_J, np.max(np.abs(J_g[k] - J_o)[nz] / np.abs(J_o)[nz]))
    print(f"[{tag}] dydt mixed-rel {worst_f:.2e} | J rel(on significant entries) {worst_J:.2e} mixed {worst_Jm:.2e} | gpu calls {tg:.2f}s")


def main():
    print("devices:", rb.lib().rmsb_device_count(), "version", rb.lib().rmsb_version())
    print("fp64 peak TFLOP/s:", rb.lib().rmsb_fp64_peak_tflops(0, 1 << 16))
    # ---- superminimal
    do = ro.readinput(str(FX / "superminimal.rms"))
    dg = rb.readinput(str(FX / "superminimal.rms"))
    igo = ro.IdealGas(do["phase"]["Species"], do["phase"]["Reactions"])
    igg = rb.IdealGas(dg["phase"]["Species"], dg["phase"]["Reactions"], name="phase")
    ic = {"T": 1000.0, "P": 10e5, "H2": 0.67, "O2": 0.33}
    od = ro.ConstantVDomain(igo, ic)
    gd, y0, p = rb.ConstantVDomain(phase=igg, initialconds=ic)
    check_domain("superminimal ConstantV", od, gd, od.p, None)
    pp = od.p.copy(); pp[:igo.Ns] = rng.normal(0, 500.0, igo.Ns); pp[igo.Ns:] = rng.uniform(0.5, 2, igo.Nr)
    check_domain("superminimal ConstantV p!=default", od, gd, pp, pp)
    ic = {"T": 1000.0, "P": 1e5, "H2": 0.67, "O2": 0.33}
    od = ro.ConstantTPDomain(igo, ic)
    gd, y0, p = rb.ConstantTPDomain(phase=igg, initialconds=ic)
    print("  TP p parity (Gs;kfs):", relerr(p, od.p))
    check_domain("superminimal ConstantTP", od, gd, od.p, p)
    # ---- ethane / propane (falloff, PLOG, Chebyshev)
    for name, ic in (("ethane", {"T": 1200.0, "P": 2e5, "ethane": 0.3, "O2": 0.6, "Ar": 0.1}),
                     ("propane", {"T": 1100.0, "P": 3e5, "propane": 0.3, "O2": 0.6, "N2": 0.1})):
        do = ro.readinput(str(FX / f"{name}.rms"))
        dg = rb.readinput(str(FX / f"{name}.rms"))
        po = ro.IdealGas(do["phase"]["Species"], do["phase"]["Reactions"])
        pg = rb.IdealGas(dg["phase"]["Species"], dg["phase"]["Reactions"])
        ic = {k: v for k, v in ic.items() if k in ("T", "P") or k in po.spcdict}
        od = ro.ConstantVDomain(po, ic)
        gd, y0, p = rb.ConstantVDomain(phase=pg, initialconds=ic)
        check_domain(f"{name} ConstantV", od, gd, od.p, None, nstate=3)
        icp = dict(ic)
        od = ro.ConstantTPDomain(po, icp)
        gd, y0, p = rb.ConstantTPDomain(phase=pg, initialconds=icp)
        print("  TP p parity:", relerr(p, od.p))
        check_domain(f"{name} ConstantTP", od, gd, od.p, p, nstate=3)
    # ---- liquid octane ConstantTV (diffusion limited)
    do = ro.readinput(str(FX / "liquid_phase.rms"))
    dg = rb.readinput(str(FX / "liquid_phase.rms"))
    lo = ro.IdealDiluteSolution(do["phase"]["Species"], do["phase"]["Reactions"], do["Solvents"][0])
    lg = rb.IdealDiluteSolution(dg["phase"]["Species"], dg["phase"]["Reactions"], dg["Solvents"][0], name="phase")
    ic = {"T": 450.0, "V": 1.0, "octane": 6154.0, "oxygen": 4.953}
    od = ro.ConstantTVDomain(lo, ic, constantspecies=["oxygen"])
    gd, y0, p = rb.ConstantTVDomain(phase=lg, initialconds=ic, constantspecies=["oxygen"])
    print("  TV kfs parity:", relerr(gd.kfs[0], od.kfs), "krevs:", relerr(gd.krevs[0], od.krevs), "Gs:", relerr(gd.Gs[0], od.Gs))
    check_domain("liquid ConstantTV", od, gd, None, None, nstate=3)
    # ---- surface ConstantTAPhi
    do = ro.readinput(str(FX / "ch4o2cat.rms"))
    dg = rb.readinput(str(FX / "ch4o2cat.rms"))
    so = ro.IdealSurface(do["surface"]["Species"], do["surface"]["Reactions"], 2.486e-5)
    sg = rb.IdealSurface(dg["surface"]["Species"], dg["surface"]["Reactions"], 2.486e-5, name="surface")
    A = 1e5 * 8.314 * 800.0 / 1e5
    ic = {"T": 800.0, "A": A, "vacantX": 2.486e-5 * A}
    od = ro.ConstantTAPhiDomain(so, ic)
    gd, y0, p = rb.ConstantTAPhiDomain(phase=sg, initialconds=ic)
    check_domain("surface ConstantTAPhi", od, gd, None, None, nstate=3)

    # ---- solves
    do = ro.readinput(str(FX / "superminimal.rms"))
    igo = ro.IdealGas(do["phase"]["Species"], do["phase"]["Reactions"])
    ic = {"T": 1000.0, "P": 10e5, "H2": 0.67, "O2": 0.33}
    od = ro.ConstantVDomain(igo, ic)
    gd, y0, p = rb.ConstantVDomain(phase=igg, initialconds=ic)
    ts = np.exp(np.linspace(np.log(1e-15), np.log(0.1), 10000))
    for rtol, atol in ((1e-6, 1e-16), (1e-10, 1e-20)):
        t0 = time.time()
        sol = rb.solve(rb.Reactor(gd, y0, (0.0, 0.101), p=p), abstol=atol, reltol=rtol, saveat=ts)
        dt = time.time() - t0
        print(f"ConstantV solve rtol={rtol}: status {sol.status} stats {sol.stats[0]} idt {sol.idt[0]:.9g} "
              f"grid-idt {rb.idt_grid(sol):.12g} (golden 0.038384723436228063) wall {dt:.2f}s")
        so_ = oi.solve(od, (0, 0.101), p=od.p, rtol=rtol, atol=atol)
        print(f"   scipy: steps {so_.t.size - 1} nfev {so_.nfev} njev {so_.njev} nlu {so_.nlu}; "
              f"yend rel diff {relerr(sol.yend[0], so_.y[:, -1]):.2e}")
    ic = {"T": 1000.0, "P": 1e5, "H2": 0.67, "O2": 0.33}
    gd, y0, p = rb.ConstantTPDomain(phase=igg, initialconds=ic)
    sol = rb.solve(rb.Reactor(gd, y0, (0.0, 150.11094), p=p), abstol=1e-20, reltol=1e-10, saveat=[20.44002454])
    sim = rb.Simulation(sol, gd)
    print("ConstantTP x(H2,O2,H2O) at 20.44:", [sim.molefractions(s, 20.44002454) for s in ("H2", "O2", "H2O")],
          "golden 0.412883111, 0.200419093, 0.386618602", sol.stats[0])
    print("ROP H2 by rxn 1:", sim.rops("H2", 20.44002454)[0], "golden -5.304207075533127e-13")
    # ensemble timing
    n = 20000
    T0 = rng.uniform(900, 1500, n); P0 = 10 ** rng.uniform(5, 6.7, n); phi = rng.uniform(0.5, 2.0, n)
    ns = np.zeros((n, igg.Ns)); ns[:, igg.spcdict["H2"] - 1] = 2 * phi; ns[:, igg.spcdict["O2"] - 1] = 1.0
    ns[:, igg.spcdict["N2"] - 1] = 3.76 * rng.integers(0, 2, n); ns /= ns.sum(axis=1, keepdims=True)
    ens = rb.ConstantVEnsemble(igg, T0, P0, ns)
    for rep in range(2):
        t0 = time.time()
        sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-16, reltol=1e-6)
        dt = time.time() - t0
        print(f"ensemble n={n}: {dt:.3f}s -> {n / dt:.0f} solves/s; status counts {np.bincount(sol.status)}; "
              f"mean steps {sol.stats['n_steps'].mean():.0f} rhs {sol.stats['n_rhs'].mean():.0f} jac {sol.stats['n_jac'].mean():.0f} lu {sol.stats['n_lu'].mean():.0f}")
    t0 = time.time(); f = ens.dydt(); t1 = time.time(); J = ens.jacobiany(); t2 = time.time()
    print(f"dydt batch {n}: {t1 - t0:.3f}s  jac batch: {t2 - t1:.3f}s (host-buffer API incl. copies)")


if __name__ == "__main__":
    main()
