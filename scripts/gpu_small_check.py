"""Group-per-reactor kernel (rmsb_small.cu) against the warp-per-reactor kernel (flags bit 5) on the C2 sweep:
statuses, ignition delays, final states, integrator statistics and wall time.  Usage: python scripts/gpu_small_check.py [n]"""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "reactionmechanismsimulator.jl_b200")); sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as ge
ge.build_cuda()
import rmsb200 as rb
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d = rb.readinput(bench.FIXTURE)
ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
y0, V = bench.conditions(ig.names, n, bench.SEED)
ens = rb.Domain(rb._lib.DOMAIN_CONSTANT_V, ig, y0, V)
res = {}
for name, flags in (("warp", 32), ("group", 0)):
    for rep in range(2):
        t0 = time.perf_counter()
        sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), rb.RMSB200(flags=flags), abstol=1e-16, reltol=1e-6)
        dt = time.perf_counter() - t0
    res[name] = sol
    st = sol.stats
    print(f"{name:6s} n={n} {n/dt:10.0f} solves/s  status!=0: {(sol.status != 0).sum()}  steps {st['n_steps'].mean():.1f} rhs {st['n_rhs'].mean():.1f} "
          f"jac {st['n_jac'].mean():.1f} lu {st['n_lu'].mean():.1f} nfail {st['n_newton_fail'].mean():.2f} efail {st['n_err_fail'].mean():.2f}", flush=True)
a, b = res["warp"], res["group"]
ign = (a.yend[:, ig.Ns] - y0[:, ig.Ns]) > 400
print("ignited", ign.sum(), "max rel idt diff", np.nanmax(np.abs(b.idt[ign] / a.idt[ign] - 1)))
print("max rel T_end diff", np.max(np.abs(b.yend[:, ig.Ns] / a.yend[:, ig.Ns] - 1)))
print("identical yend rows", int((a.yend == b.yend).all(axis=1).sum()), "of", n)
# tight tolerance + saved outputs
tq = np.array([1e-4, 1e-3, 1e-2, 0.05, 0.1])
for name, flags in (("warp", 32), ("group", 0)):
    sol = rb.solve(rb.Reactor(rb.Domain(rb._lib.DOMAIN_CONSTANT_V, ig, y0[:256], V[:256]), None, (0.0, 0.1)), rb.RMSB200(flags=flags), abstol=1e-20, reltol=1e-10, saveat=tq)
    res[name + "_t"] = sol
    print(name, "tight: status", (sol.status != 0).sum(), "steps", sol.stats["n_steps"].mean())
a, b = res["warp_t"], res["group_t"]
big = np.abs(a.u) > 1e-12 * np.max(np.abs(a.u[:, :, :ig.Ns]))
print("tight trajectories max rel diff", np.max(np.abs(a.u - b.u)[big] / np.abs(a.u)[big]))
