"""Secondary throughput numbers for the BASELINE.json configs other than the headline one (C2 is bench.py):
C1 superminimal ConstantTP ensemble, C3 liquid octane (35 species) and a synthetic 1000-species ConstantTV mechanism,
C4 coupled gas + catalyst, C5 forward sensitivities.  Host-buffer C-ABI calls, wall clock around the call
(copies included), second of two runs.  Usage (GPU box):  python scripts/bench_configs.py > gpurun_out/configs.txt"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import cases  # noqa: E402
import synth  # noqa: E402
import rmsb200 as rb  # noqa: E402


def timed(fn, reps=2):
    for _ in range(reps):
        t0 = time.time()
        out = fn()
        dt = time.time() - t0
    return out, dt


def report(tag, n, dt, sol, extra=""):
    st = sol.stats
    print(f"{tag}: n={n} {n / dt:.1f} solves/s  ok {np.mean(sol.status == 0):.4f}  mean steps {st['n_steps'].mean():.0f} rhs {st['n_rhs'].mean():.0f} "
          f"jac {st['n_jac'].mean():.1f} lu {st['n_lu'].mean():.0f} {extra}", flush=True)


rng = np.random.default_rng(11)
# ---- C1 / C5: superminimal ConstantTP
gdom, y0, p = cases.product_domain("superminimal_TP")
ph = gdom.phase
n = 100000
T = rng.uniform(950, 1250, n); P = rng.uniform(0.5e5, 5e5, n)
ns = np.zeros((n, ph.Ns)); ns[:, ph.names.index("H2")] = rng.uniform(0.3, 0.7, n); ns[:, ph.names.index("O2")] = rng.uniform(0.2, 0.4, n)
ens = rb.ConstantTPEnsemble(ph, T, P, ns)
sol, dt = timed(lambda: rb.solve(rb.Reactor(ens, None, (0.0, 1.0)).ode, abstol=1e-16, reltol=1e-6))
report("C1 superminimal ConstantTP, t_end 1 s, rtol 1e-6", n, dt, sol)
n5 = 20000
ens5 = rb.ConstantTPEnsemble(ph, T[:n5], P[:n5], ns[:n5])
sol, dt = timed(lambda: rb.solve(rb.Reactor(ens5, None, (0.0, 1.0), forwardsensitivities=True).ode, abstol=1e-21, reltol=1e-7))
report("C5 forward sensitivities (29 parameters), reference tolerances 1e-21 / 1e-7", n5, dt, sol, f"(= {n5 * 29 / dt:.0f} parameter-sensitivity trajectories/s)")
# ---- C3: liquid octane oxidation, 35 species (the reference's fixture) and the synthetic 1000-species mechanism
gdom, y0, p = cases.product_domain("liquid_TV")
n = 20000
T = rng.uniform(420, 480, n)
y0s = np.repeat(gdom.y0, n, axis=0) * rng.uniform(0.8, 1.2, (n, 1))
ens = rb.Domain(gdom.kind, gdom.phase, y0s, np.stack([T, np.ones(n)], axis=1), constantspecies=["oxygen"])
sol, dt = timed(lambda: rb.solve(rb.Reactor(ens, None, (0.0, 1.4e5)).ode, abstol=1e-16, reltol=1e-6))
report("C3 liquid octane (35 species, diffusion limited), t_end 1.4e5 s", n, dt, sol)
for nsp, nre in ((300, 1184), (1000, 592)):
    path = f"/tmp/synth{nsp}.rms"
    synth.write_synthetic_mechanism(path, ns=nsp, seed=7)
    d = rb.readinput(path)
    sph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase", diffusionlimited=False)
    sdom, sy0, sp = rb.ConstantTVDomain(phase=sph, initialconds={"T": 650.0, "V": 1.0, **synth.initial_moles(nsp)})
    T = rng.uniform(600, 700, nre)
    y0s = np.repeat(sdom.y0, nre, axis=0) * rng.uniform(0.8, 1.2, (nre, 1))
    ens = rb.Domain(sdom.kind, sph, y0s, np.stack([T, np.ones(nre)], axis=1))
    sol, dt = timed(lambda: rb.solve(rb.Reactor(ens, None, (0.0, 1.0)).ode, abstol=1e-12, reltol=1e-6))
    report(f"C3-large synthetic {nsp} species / {sph.Nr} reactions ConstantTV (block per reactor, sparse LU)", nre, dt, sol)
# ---- C4: coupled gas + catalyst
from test_gpu_coupled import build  # noqa: E402
react, y0, p, doms, inter = build()
cph = inter.phase
N1, N2 = doms[0].phase.Ns, doms[1].phase.Ns
n = 10000
T = rng.uniform(700, 1000, n); x = rng.uniform(0.05, 0.2, n)
nsg = np.zeros((n, N1)); nss = np.zeros((n, N2))
gn, sn = doms[0].phase.names, doms[1].phase.names
nsg[:, gn.index("O2")], nsg[:, gn.index("N2")], nsg[:, gn.index("CH4")] = 0.2, 0.7, x
Pp = np.full(n, 1.0e5)
V = nsg.sum(axis=1) * 8.314 * T / Pp
A = 1.0e5 * V
nss[:, sn.index("vacantX")] = 2.486e-5 * A
ens = rb.CoupledEnsemble(cph, T, Pp, A, nsg, nss)
sol, dt = timed(lambda: rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-20, reltol=1e-6))
report("C4 gas (12) + Pt surface (29) + interface, Ny 42, 10 000 conditions, t_end 0.1 s", n, dt, sol)
