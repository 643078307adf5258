"""The N > 1 path on CPU: world_size-2 gloo run of the shard + gather logic (rmsb200/ensemble.py).  The per-rank
solver is injected: here the C++ CPU oracle stands in for the CUDA engine, so the test also checks that a sharded
run reproduces the single-process run reactor by reactor."""
import os
import sys

import numpy as np
import pytest

import cases


def test_shard_bounds_cover_and_balance():
    from rmsb200.ensemble import shard_bounds
    for n in (0, 1, 7, 8, 9, 100000, 100003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_shards_are_a_balanced_permutation():
    from rmsb200.ensemble import balanced_shards
    rng = np.random.default_rng(0)
    work = rng.lognormal(6.0, 0.6, 100003)
    for world in (1, 2, 4, 8):
        parts = balanced_shards(work, world, 64)
        allidx = np.concatenate(parts)
        assert np.array_equal(np.sort(allidx), np.arange(len(work)))
        tot = np.array([work[p].sum() for p in parts])
        assert tot.max() / tot.mean() < 1.02                       # same mix on every rank (the excess is the single heaviest chunk)
        for p in parts:
            assert np.all(np.diff(work[p]) <= 0)                   # each rank's queue runs heaviest first
    assert len(balanced_shards(np.ones(3), 8, 64)) == 8             # more ranks than chunks: empty shards are fine


def _worker(rank, world, port, n, outfile):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (str(cases.ROOT), str(cases.ROOT / "reactionmechanismsimulator.jl_b200"), str(cases.ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import rmsb200 as rb
    from oracle import cpu_oracle as co
    from rmsb200.ensemble import solve_sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gdom_phase = rb.readinput(cases.fixture("superminimal"))
        ig = rb.IdealGas(gdom_phase["phase"]["Species"], gdom_phase["phase"]["Reactions"])
        rng = np.random.default_rng(5)
        T0, P0 = rng.uniform(1000, 1400, n), 10 ** rng.uniform(5, 6, n)
        ns = np.zeros((n, ig.Ns)); ns[:, ig.spcdict["H2"] - 1] = 0.67; ns[:, ig.spcdict["O2"] - 1] = 0.33
        ens = rb.ConstantVEnsemble(ig, T0, P0, ns)
        cm = co.CpuMech(cases.oracle_domain("superminimal_V").phase)

        def solver(d, te, atol, rtol):
            r = cm.solve("ConstantV", d.y0, d.cond, te, rtol, atol, nthreads=1)
            return r["idt"], r["status"], r["yend"], r["stats"]["n_steps"]

        res = solve_sharded(ens, 0.05, abstol=1e-16, reltol=1e-6, solver=solver)
        if rank == 0:
            ref = solver(ens, np.full(n, 0.05), 1e-16, 1e-6)
            np.savez(outfile, idt=res.idt, status=res.status, yend=res.yend, nst=res.n_steps,
                     ridt=ref[0], rstatus=ref[1], ryend=ref[2], rnst=ref[3])
        else:
            assert res.idt is None and (res.local_hi - res.local_lo) in (n // world, n // world + 1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [9, 2])
def test_two_rank_gloo_shard_and_gather(tmp_path, n):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npz")
    port = 29500 + (os.getpid() % 2000) + n
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    d = np.load(out)
    assert d["idt"].shape == (n,) and d["yend"].shape == (n, 14)
    assert np.array_equal(d["status"], d["rstatus"]) and np.array_equal(d["nst"], d["rnst"])
    assert np.array_equal(d["idt"], d["ridt"], equal_nan=True) and np.array_equal(d["yend"], d["ryend"])
