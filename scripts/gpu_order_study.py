"""Work-order study for the headline sweep: per-reactor cost (n_rhs) against the initial conditions, and the makespan of list
scheduling on the resident lane groups for several orders (as given, heaviest first, by T0, ...).  Usage: python scripts/gpu_order_study.py"""
import sys
import heapq
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200"):
    sys.path.insert(0, str(p))
import bench  # noqa: E402
import rmsb200 as rb  # noqa: E402

d = rb.readinput(bench.FIXTURE)
ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
n = 100000
y0, V = bench.conditions(ig.names, n, bench.SEED)
ens = rb.Domain(rb.DOMAIN_CONSTANT_V, ig, y0, V)
sol = rb.solve(rb.Reactor(ens, None, (0.0, bench.T_END)).ode, abstol=bench.ATOL, reltol=bench.RTOL)
cost = sol.stats["n_rhs"].astype(float) + 4.0 * sol.stats["n_lu"] + 8.0 * sol.stats["n_jac"]
T0, P0 = y0[:, ig.Ns], y0[:, ig.Ns + 1]
xH2 = y0[:, ig.names.index("H2")]
xN2 = y0[:, ig.names.index("N2")]
print("cost mean", cost.mean(), "max", cost.max(), "p99", np.percentile(cost, 99), "min", cost.min())
for name, v in (("T0", T0), ("logP0", np.log10(P0)), ("xH2", xH2), ("xN2", xN2), ("ignited", np.isfinite(sol.idt) & (sol.yend[:, ig.Ns] > T0 + 400))):
    print(f"corr(cost, {name}) = {np.corrcoef(cost, v.astype(float))[0, 1]:.3f}")


def makespan(order, workers=4736):
    h = [0.0] * workers
    heapq.heapify(h)
    for i in order:
        t = heapq.heappop(h)
        heapq.heappush(h, t + cost[i])
    return max(h)


ideal = cost.sum() / 4736
print("ideal (sum/workers)", ideal)
for name, order in (("as given", np.arange(n)), ("heaviest first (oracle)", np.argsort(-cost)), ("T0 descending", np.argsort(-T0)), ("T0 ascending", np.argsort(T0)),
                    ("T0*logP descending", np.argsort(-(T0 * np.log10(P0)))), ("1000/T0 - 0.5 logP ascending", np.argsort(1000.0 / T0 - 0.5 * np.log10(P0)))):
    print(f"{name:34s} makespan / ideal = {makespan(order) / ideal:.3f}")
# a fitted predictor: quadratic in (1000/T0, log10 P0, xH2, xN2)
X = np.stack([np.ones(n), 1000 / T0, np.log10(P0), xH2, xN2, (1000 / T0) ** 2, np.log10(P0) ** 2, (1000 / T0) * np.log10(P0)], axis=1)
coef, *_ = np.linalg.lstsq(X[:5000], cost[:5000], rcond=None)
pred = X @ coef
print("fitted quadratic predictor: corr", np.corrcoef(pred, cost)[0, 1], "makespan / ideal", makespan(np.argsort(-pred)) / ideal)

# ---- measured kernel time for several input orders (the two reactors of a warp run in lockstep: similar neighbours waste less)
def run(order, tag):
    e = rb.Domain(rb.DOMAIN_CONSTANT_V, ig, np.ascontiguousarray(y0[order]), np.ascontiguousarray(V[order]))
    best = 1e30
    for _ in range(3):
        s = rb.solve(rb.Reactor(e, None, (0.0, bench.T_END)).ode, abstol=bench.ATOL, reltol=bench.RTOL)
        best = min(best, s.kernel_ms)
    print(f"{tag:40s} {n / best * 1e3:10.0f} solves/s (kernel)")


def morton(*cols, bits=10):
    q = [np.clip(((c - c.min()) / (np.ptp(c) + 1e-300) * ((1 << bits) - 1)).astype(np.uint64), 0, (1 << bits) - 1) for c in cols]
    key = np.zeros(n, dtype=np.uint64)
    for b in range(bits):
        for k, c in enumerate(q):
            key |= ((c >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * len(q) + k)
    return key


run(np.arange(n), "as given")
run(np.argsort(-cost), "cost descending (pilot knowledge)")
run(np.argsort(-T0), "T0 descending")
run(np.argsort(morton(T0, np.log10(P0))), "Morton(T0, logP0)")
run(np.argsort(morton(T0, np.log10(P0), xH2)), "Morton(T0, logP0, xH2)")
run(np.lexsort((np.log10(P0), T0, xN2 > 0)), "lexsort(dilution, T0, logP0)")
run(np.argsort(morton(T0, np.log10(P0), xH2, xN2)), "Morton(T0, logP0, xH2, xN2)")
