"""Multi-domain reactors WITHOUT interfaces (src/Reactor.jl:99-258 with `interfaces = []`; src/TestReactors.jl:470-531): several
domains integrated as one system.  With nothing coupling them the system is block diagonal, so every domain is solved by its own
batched call (any mix of domain kinds) and the results are presented in the reference's combined layout:
    y = [species of domain 1; species of domain 2; ...; thermodynamic variables, filled from the END backwards domain by domain]
(src/Reactor.jl:106-133: the first thermodynamic variable of domain 1 is the last entry, its second the one before, then domain 2's).
The parameter vector is the concatenation of the domains' p, one block per distinct phase (:135-165)."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


class MultiDomainReactor:
    def __init__(self, domains: Sequence, y0s: Sequence, tspan, ps: Optional[Sequence] = None, forwardsensitivities: bool = False):
        from .reactor import Reactor
        if forwardsensitivities:
            raise NotImplementedError("forward sensitivities of multi-domain reactors: solve the domains' own reactors with forwardsensitivities=True")
        self.domains = list(domains)
        self.tspan = (float(tspan[0]), tspan[1])
        ps = [None] * len(self.domains) if ps is None else list(ps)
        self.reactors: List = [Reactor(d, y, tspan, p=p) for d, y, p in zip(self.domains, y0s, ps)]
        self.Ns = [d.phase.Ns for d in self.domains]
        self.Nth = [d.Ny - d.phase.Ns for d in self.domains]
        self.Ny = sum(d.Ny for d in self.domains)
        # parameter vector: one block per distinct phase, in order of first appearance
        blocks, seen = [], []
        for d, p in zip(self.domains, ps):
            if any(d.phase is q for q in seen) or p is None:
                continue
            seen.append(d.phase)
            blocks.append(np.ravel(np.asarray(p, dtype=float)))
        self.p = np.concatenate(blocks) if blocks else None
        self.y0 = self.combine([np.atleast_2d(r.y0)[0] for r in self.reactors])
        self.forwardsensitivities = False

    @property
    def ode(self):
        return self

    def combine(self, ys: Sequence[np.ndarray]) -> np.ndarray:
        """Per-domain states -> the reference's combined state vector."""
        out = np.empty(self.Ny)
        k, n = 0, self.Ny - 1
        for y, ns, nth in zip(ys, self.Ns, self.Nth):
            out[k:k + ns] = y[:ns]
            k += ns
            for m in range(nth):
                out[n] = y[ns + m]
                n -= 1
        return out


class MultiDomainSolution:
    """`sol(t)` returns the combined state; `sols[i]` are the domains' own solutions."""

    def __init__(self, reactor: MultiDomainReactor, sols: List):
        self.reactor, self.sols = reactor, sols
        self.t = sols[0].t
        self.retcode = ["Success" if all(s.retcode[0] == "Success" for s in sols) else "Failure"]
        self.status = np.array([max(int(s.status[0]) for s in sols)], dtype=np.int32)

    def __call__(self, t, r: int = 0):
        return self.reactor.combine([s(t, r) for s in self.sols])

    @property
    def yend(self):
        return self.reactor.combine([s.yend[0] for s in self.sols])[None]


def solve_multidomain(react: MultiDomainReactor, alg=None, **kw) -> MultiDomainSolution:
    from .reactor import solve
    return MultiDomainSolution(react, [solve(r, alg, **kw) for r in react.reactors])


class MultiDomainSimulation:
    """SystemSimulation(sol, domains, [], p) for domains without interfaces: `sims[i]` is the Simulation of domain i (src/Simulation.jl:92-134)."""

    def __init__(self, sol: MultiDomainSolution, domains=None, interfaces=(), p=None):
        from .simulation import Simulation
        self.sol = sol
        self.domains = list(domains) if domains is not None else sol.reactor.domains
        self.sims = [Simulation(s, d) for s, d in zip(sol.sols, self.domains)]
        self.names = [nm for d in self.domains for nm in d.phase.names]
        self.p = sol.reactor.p if p is None else p
