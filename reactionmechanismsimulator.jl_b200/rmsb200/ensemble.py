"""Multi-GPU ensembles: reactors are independent, so the condition index range is cut into contiguous
blocks, one per rank (one process per GPU, `torch.distributed`); mechanism tables are replicated; the
only exchange is the gather of ignition delays / status / final states (and saved trajectories) on rank 0
-- NCCL over NVLink on GPUs, gloo in the CPU tests.

The reference's only parallel construct is a thread loop with one deep copy of the reactor per thread
(src/ThreadedSensitivities.jl:40-55); this is its many-reactor, many-device counterpart.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np

from . import _lib as L
from .domain import Domain


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first n % world ranks get one extra reactor."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_shards(work, world: int, chunk: int = 64):
    """Strong-scaling split: reactors sorted by a work estimate (e.g. n_rhs of a pilot solve at loose tolerance, or of
    the previous sweep), heaviest first, then dealt round-robin (in snake order) in chunks so every rank gets the same mix and starts with its
    heaviest reactors (longest-processing-time order inside each rank's work queue).  Returns one index array per rank; together
    they are a permutation of range(len(work))."""
    work = np.asarray(work)
    order = np.argsort(-work, kind="stable")
    nchunks = (len(order) + chunk - 1) // chunk
    parts = [[] for _ in range(world)]
    for c in range(nchunks):                       # snake order: 0..W-1, W-1..0, ... so no rank always gets the heavier chunk of a round
        rnd, pos = divmod(c, world)
        parts[pos if rnd % 2 == 0 else world - 1 - pos].append(order[c * chunk:(c + 1) * chunk])
    return [np.concatenate(p) if p else np.empty(0, dtype=order.dtype) for p in parts]


def shard_domain(dom: Domain, world: int, rank: int) -> Domain:
    """The rank's slice of an ensemble domain (same phase handle semantics, own device tables)."""
    lo, hi = shard_bounds(dom.n, world, rank)
    sub = Domain(dom.kind, dom.phase, dom.y0[lo:hi], dom.cond[lo:hi])
    sub.constantspeciesinds = dom.constantspeciesinds
    sub.p, sub._p_default = dom.p, dom._p_default
    if dom.p is not None and np.ndim(dom.p) == 2:
        sub.p = dom.p[lo:hi]
    return sub


@dataclass
class ShardedResult:
    """Rank 0 holds the gathered arrays for the whole ensemble; other ranks hold None."""
    idt: Optional[np.ndarray]
    status: Optional[np.ndarray]
    yend: Optional[np.ndarray]
    n_steps: Optional[np.ndarray]
    local_lo: int
    local_hi: int


def solve_sharded(dom: Domain, t_end, *, abstol: float, reltol: float, solver: Optional[Callable] = None, group=None,
                  device=None) -> ShardedResult:
    """Solve this rank's block of `dom` and gather (idt, status, yend, n_steps) on rank 0.

    `solver(sub_domain, t_end_block, abstol, reltol) -> (idt, status, yend, n_steps)`; the default runs the CUDA
    engine (`rmsb200.solve`).  Must be called by every rank of the process group (or without one: world = 1).
    """
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_bounds(dom.n, world, rank)
    sub = shard_domain(dom, world, rank)
    te = np.broadcast_to(np.asarray(t_end, dtype=float), (dom.n,))[lo:hi]
    if solver is None:
        from .reactor import Reactor, solve

        def solver(d, te_, atol, rtol):
            sol = solve(Reactor(d, None, (0.0, float(np.max(te_)) if te_.size else 0.0)), abstol=atol, reltol=rtol, t_end=te_)
            return sol.idt, sol.status, sol.yend, sol.stats["n_steps"]
    idt, status, yend, nst = solver(sub, te, abstol, reltol) if hi > lo else (np.empty(0), np.empty(0, np.int32),
                                                                                np.empty((0, dom.Ny)), np.empty(0, np.int32))
    if world == 1:
        return ShardedResult(np.asarray(idt), np.asarray(status), np.asarray(yend), np.asarray(nst), lo, hi)
    backend = dist.get_backend(group)
    dev = torch.device("cpu") if backend == "gloo" else (device or torch.device("cuda", torch.cuda.current_device()))
    # pad every block to the largest block so a single gather per array suffices
    nmax = shard_bounds(dom.n, world, 0)[1]
    Ny = dom.Ny

    def padded(a, shape, dtype):
        t = torch.zeros(shape, dtype=dtype)
        t[: hi - lo] = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
        return t.to(dev)

    payload = [padded(idt, (nmax,), torch.float64), padded(status, (nmax,), torch.int32),
               padded(yend, (nmax, Ny), torch.float64), padded(nst, (nmax,), torch.int32)]
    out = []
    for t in payload:
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, bufs, dst=0, group=group)
        out.append(bufs)
    if rank != 0:
        return ShardedResult(None, None, None, None, lo, hi)
    res = []
    for bufs in out:
        parts = []
        for r, b in enumerate(bufs):
            l, h = shard_bounds(dom.n, world, r)
            parts.append(b[: h - l].cpu().numpy())
        res.append(np.concatenate(parts, axis=0))
    return ShardedResult(res[0], res[1], res[2], res[3], lo, hi)
