"""Integrator-flag sweep: accuracy (vs the C++ CPU oracle at rtol 1e-10) and speed at the throughput tolerances."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "reactionmechanismsimulator.jl_b200")
import numpy as np, cases, bench
import rmsb200 as rb
from oracle import cpu_oracle as co
gdom, _, _ = cases.product_domain("superminimal_V")
ig = gdom.phase
y0, V = bench.conditions(ig.names, 50000, 20261019)
ens = rb.Domain(gdom.kind, ig, y0, V)
nref = 2000
ref = co.CpuMech(cases.oracle_domain("superminimal_V").phase).solve("ConstantV", y0[:nref], V[:nref], 0.1, 1e-10, 1e-20)
ign = (ref["yend"][:, 12] - y0[:nref, 12]) > 400
for flags in [int(x, 0) for x in sys.argv[1:]]:
    alg = rb.RMSB200(flags=flags)
    rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), alg, abstol=1e-16, reltol=1e-6)
    t0 = time.time(); sol = rb.solve(rb.Reactor(ens, None, (0.0, 0.1)), alg, abstol=1e-16, reltol=1e-6); dt = time.time() - t0
    st = sol.stats
    eidt = np.nanmax(np.abs(sol.idt[:nref][ign] / ref["idt"][ign] - 1))
    eT = np.max(np.abs(sol.yend[:nref, 12] / ref["yend"][:, 12] - 1))
    print(f"flags=0x{flags:x}: {50000/dt:8.0f} solves/s steps {st['n_steps'].mean():6.1f} rhs {st['n_rhs'].mean():7.1f} lu {st['n_lu'].mean():6.1f} jac {st['n_jac'].mean():5.1f} "
          f"nfail {st['n_newton_fail'].mean():5.1f} efail {st['n_err_fail'].mean():5.1f} | idt err {eidt:.2e} Tend err {eT:.2e} ok {int((sol.status==0).all())}")
