import sys
sys.path.insert(0, "/root/repo/reactionmechanismsimulator.jl_b200"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import numpy as np, cases, rmsb200 as rb
gdom, y0, p = cases.product_domain("propane_V")
for rt in (1e-6, 1e-8):
    sol = rb.solve(rb.Reactor(gdom, y0, (0.0, 2.0e-3), p=p), abstol=1e-18, reltol=rt)
    print(rt, sol.retcode, sol.stats[0], sol.yend[0, -2:])
