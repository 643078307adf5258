"""Throughput of the BASELINE.json configs other than the headline one (C2 is bench.py's timed region): C1 superminimal
ConstantTP ensemble, C3 liquid octane (35 species, the reference's fixture) and C3 at its stated size (the octane template
replicated over 36 fuel families: 1015 species / 11.6 k reactions, tests/synth.py), C4 coupled gas + catalyst, C5 forward
sensitivities.  Every config goes through the public host-buffer call (`rb.solve` -> rmsb_solve_batch / rmsb_sens_batch):
`e2e` is wall clock around the call (copies included), `value` the device time of its kernels (CUDA events inside the library,
rmsb_last_kernel_ms), second of two runs.

bench.py imports `run_all` for the `others` block of its JSON line;  python scripts/bench_configs.py prints the same block."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

STATUS = ["ok", "max_steps", "step_too_small", "nan"]


def _timed(fn, reps=2):
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
    return out, dt


def _record(workload, n, dt, sol, extra=None):
    st = sol.stats
    rec = {"workload": workload, "conditions": int(n), "value": n / (sol.kernel_ms * 1e-3) if sol.kernel_ms > 0 else None,
           "e2e": n / dt, "unit": "solves/s", "kernel_ms": sol.kernel_ms, "wall_ms": dt * 1e3,
           "status": {STATUS[k]: int(c) for k, c in enumerate(np.bincount(sol.status, minlength=4)[:4])},
           "mean_per_solve": {k: float(st[k].mean()) for k in ("n_steps", "n_rhs", "n_jac", "n_lu", "n_newton_fail", "n_err_fail")},
           "steps_per_s": float(st["n_steps"].sum()) / (sol.kernel_ms * 1e-3) if sol.kernel_ms > 0 else None}
    if extra:
        rec.update(extra)
    return rec


def c1_c5(rb, cases, sizes, device=0):
    rng = np.random.default_rng(11)
    gdom, y0, p = cases.product_domain("superminimal_TP")
    ph = gdom.phase
    n = sizes["C1"]
    T = rng.uniform(950, 1250, n); P = rng.uniform(0.5e5, 5e5, n)
    ns = np.zeros((n, ph.Ns)); ns[:, ph.names.index("H2")] = rng.uniform(0.3, 0.7, n); ns[:, ph.names.index("O2")] = rng.uniform(0.2, 0.4, n)
    ens = rb.ConstantTPEnsemble(ph, T, P, ns)
    sol, dt = _timed(lambda: rb.solve(rb.Reactor(ens, None, (0.0, 1.0)).ode, abstol=1e-16, reltol=1e-6))
    out = {"C1": _record("configs[0] as an ensemble: H2/O2 (superminimal) ConstantTP, T 950-1250 K, t_end 1 s, rtol 1e-6 / atol 1e-16", n, dt, sol)}
    n5 = sizes["C5"]
    ens5 = rb.ConstantTPEnsemble(ph, T[:n5], P[:n5], ns[:n5])
    sol, dt = _timed(lambda: rb.solve(rb.Reactor(ens5, None, (0.0, 1.0), forwardsensitivities=True).ode, abstol=1e-21, reltol=1e-7))
    out["C5"] = _record("configs[4]: forward sensitivities of the H2/O2 ConstantTP ensemble, all 29 parameters, rtol 1e-7 / atol 1e-21 "
                        "(src/TestReactors.jl:232-236)", n5, dt, sol, {"parameter_trajectories_per_s": n5 * 29 / (sol.kernel_ms * 1e-3)})
    return out


def c3_small(rb, cases, sizes):
    rng = np.random.default_rng(12)
    gdom, y0, p = cases.product_domain("liquid_TV")
    n = sizes["C3_35"]
    T = rng.uniform(420, 480, n)
    y0s = np.repeat(gdom.y0, n, axis=0) * rng.uniform(0.8, 1.2, (n, 1))
    ens = rb.Domain(gdom.kind, gdom.phase, y0s, np.stack([T, np.ones(n)], axis=1), constantspecies=["oxygen"])
    sol, dt = _timed(lambda: rb.solve(rb.Reactor(ens, None, (0.0, 1.4e5)).ode, abstol=1e-16, reltol=1e-6))
    return {"C3_35": _record("configs[2] on the reference's fixture: liquid octane oxidation ConstantTV, 35 species / 380 reactions, diffusion "
                             "limited, T 420-480 K, t_end 1.4e5 s (src/TestReactors.jl:18-22)", n, dt, sol)}


def c3_large(rb, sizes, families=36):
    import synth
    rng = np.random.default_rng(13)
    path = f"/tmp/rmsb_octane_families_{families}.rms"
    info = synth.write_octane_families(path, families=families)
    d = rb.readinput(path)
    ph = rb.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], name="phase")
    gdom, y0, p = rb.ConstantTVDomain(phase=ph, initialconds={"T": 450.0, "V": 1.0, **synth.octane_family_moles(families)}, constantspecies=["oxygen"])
    n = sizes["C3_large"]
    T = rng.uniform(400.0, 500.0, n)
    y0s = np.repeat(gdom.y0, n, axis=0) * rng.uniform(0.8, 1.2, (n, 1))
    y0s[:, ph.names.index("oxygen")] = 4.953 * rng.uniform(0.5, 2.0, n)
    ens = rb.Domain(gdom.kind, ph, y0s, np.stack([T, np.ones(n)], axis=1), constantspecies=["oxygen"])
    sol, dt = _timed(lambda: rb.solve(rb.Reactor(ens, None, (0.0, 140000.01)).ode, abstol=1e-16, reltol=1e-6))
    li = np.zeros(8, dtype=np.int64)
    rb.lib().rmsb_large_info(ph.handle, li.ctypes.data)
    return {"C3_large": _record(f"configs[2] at its stated size: octane template replicated over {families} fuel families sharing the O2/HO2/OH pool "
                                f"(SURVEY 8d recipe), {info['ns']} species / {info['nr']} reactions, ConstantTV diffusion limited, T 400-500 K, [O2] x0.5-2, "
                                "t_end 1.4e5 s, block per reactor + fixed-pattern sparse LU", n, dt, sol,
                                {"nnz_LU": int(li[1]), "levels": int(li[2]), "dense_core": int(li[0] - li[6]), "factor_fma": int(li[5])})}


def c4(rb, sizes):
    from test_gpu_coupled import build
    rng = np.random.default_rng(14)
    react, y0, p, doms, inter = build()
    cph = inter.phase
    N1, N2 = doms[0].phase.Ns, doms[1].phase.Ns
    n = sizes["C4"]
    T = rng.uniform(700, 1000, n); x = rng.uniform(0.05, 0.2, n)
    nsg = np.zeros((n, N1)); nss = np.zeros((n, N2))
    gn, sn = doms[0].phase.names, doms[1].phase.names
    nsg[:, gn.index("O2")], nsg[:, gn.index("N2")], nsg[:, gn.index("CH4")] = 0.2, 0.7, x
    Pp = np.full(n, 1.0e5)
    V = nsg.sum(axis=1) * 8.314 * T / Pp
    A = 1.0e5 * V
    nss[:, sn.index("vacantX")] = 2.486e-5 * A
    ens = rb.CoupledEnsemble(cph, T, Pp, A, nsg, nss)
    sol, dt = _timed(lambda: rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), abstol=1e-20, reltol=1e-6))
    return {"C4": _record("configs[3]: CH4/O2 gas (ConstantTP, 12 species) + Pt surface (ConstantTAPhi, 29 species) + reactive interface, Ny 42, "
                          "T 700-1000 K, t_end 0.1 s, rtol 1e-6 / atol 1e-20 (src/TestReactors.jl:545-557)", n, dt, sol)}


DEFAULT_SIZES = {"C1": 100000, "C5": 10000, "C3_35": 20000, "C3_large": 1184, "C4": 10000}


def run_all(sizes=None):
    import cases
    import rmsb200 as rb
    sizes = {**DEFAULT_SIZES, **(sizes or {})}
    out = {}
    for fn in (lambda: c1_c5(rb, cases, sizes), lambda: c3_small(rb, cases, sizes), lambda: c3_large(rb, sizes), lambda: c4(rb, sizes)):
        try:
            out.update(fn())
        except Exception as exc:          # one config failing must not take the headline number down with it
            out[f"error_{len(out)}"] = repr(exc)
    return out


if __name__ == "__main__":
    print(json.dumps(run_all(), indent=1))
