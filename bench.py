#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched stiff-kinetics hot path (BASELINE.json).

Metric : ensemble reactor solves/sec to t_end (IDT sweep); dydt + Jacobian evals/sec reported beside it.
Workload (configs[1]): H2/O2 (superminimal mechanism) ConstantVDomain ignition-delay ensemble, 100 000
(T0, P0, phi) conditions per GPU, T0 in [900,1500] K, log10 P0 in [5,6.7], phi in [0.5,2], N2 dilution 0/3.76,
t_end = 0.1 s, rtol 1e-6 / atol 1e-16.  One "step" = the whole ensemble solved once.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n-per-gpu M] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); ranks solve disjoint condition sets (weak scaling), the only
collective is the NCCL gather of ignition delays / status to rank 0, inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "reactionmechanismsimulator.jl_b200"))

FIXTURE = str(ROOT / "tests" / "fixtures" / "superminimal.rms")
METRIC = "ensemble reactor solves/sec to t_end (IDT sweep)"
UNIT = "solves/s"
RTOL, ATOL, T_END = 1e-6, 1e-16, 0.1
SEED = 20261019


def conditions(names, n, seed):
    """SURVEY 8(d) C2 sweep -> y0 [n][14], V [n][1]."""
    rng = np.random.default_rng(seed)
    T0, P0, phi = rng.uniform(900, 1500, n), 10 ** rng.uniform(5, 6.7, n), rng.uniform(0.5, 2.0, n)
    ns = np.zeros((n, len(names)))
    ns[:, names.index("H2")] = 2 * phi
    ns[:, names.index("O2")] = 1.0
    ns[:, names.index("N2")] = 3.76 * rng.integers(0, 2, n)
    ns /= ns.sum(axis=1, keepdims=True)
    V = ns.sum(axis=1) * 8.314 * T0 / P0              # src/Domain.jl:143-145
    return np.ascontiguousarray(np.concatenate([ns, T0[:, None], P0[:, None]], axis=1)), np.ascontiguousarray(V[:, None])


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _issue_roofline(mean_steps, n, t_kernel, clocks, device):
    """The bound that actually limits the solve kernel: warp-instruction issue.  Warp instructions per accepted step come from the
    committed ncu capture of the bench workload (smsp__inst_executed.sum 1.264e11 / 100 000 reactors / 391.1 steps,
    profiles/r02_small_solve_kernel_ncu_k.txt); the peak is
    SMs x 4 schedulers x one instruction per cycle at the SM clock sampled during the timed region."""
    try:
        import torch
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz")
        per_step = 3232.0
        ach = per_step * mean_steps * n / t_kernel
        peak = sms * 4 * float(mhz) * 1e6
        return {"achieved_ginst_s": ach / 1e9, "peak_ginst_s": peak / 1e9, "frac": ach / peak, "warp_instructions_per_step": per_step,
                "source": "profiles/r02_small_solve_kernel_ncu_k.txt (ncu --set full of this workload)"}
    except Exception as exc:
        return {"achieved_ginst_s": None, "peak_ginst_s": None, "frac": None, "source": f"unavailable: {exc}"}


def _ncu_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum of one solve-kernel launch from the committed `ncu --set full` capture
    (profiles/r02_solve_kernel_traffic.json); only valid for the launch size it was captured at, otherwise None."""
    p = ROOT / "profiles" / "r02_solve_kernel_traffic.json"
    try:
        d = json.loads(p.read_text())
        if int(d["conditions_per_launch"]) == int(n):
            return int(d["dram__bytes_read.sum"]) + int(d["dram__bytes_write.sum"])
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path.  RMS is a Julia package (no julia in this
# image, nothing to compile into oracle/_ref), so this is the C++ port oracle/cpu/rms_cpu.cpp on all host cores.
# ------------------------------------------------------------------------------------------------------
CPU_BUILD_NOTE = ("C++/OpenMP port of the RMS CPU path (oracle/cpu/rms_cpu.cpp) built for timing with g++ -O3 -march=native -ffast-math "
                  "-fno-finite-math-only on this host (the strict -O2 -ffp-contract=off build is the checker only); RMS itself is Julia "
                  "and cannot run in this image")


def cpu_port(fast=True):
    from oracle import cpu_oracle as co
    from oracle import rms_oracle as ro
    co.select_build("fast" if fast else "strict")
    d = ro.readinput(FIXTURE)
    ph = ro.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"])
    return co, co.CpuMech(ph), ph


def cpu_solve_rate(cm, names, n, seed, nthreads):
    y0, V = conditions(names, n, seed)
    t0 = time.perf_counter()
    r = cm.solve("ConstantV", y0, V, T_END, RTOL, ATOL, nthreads=nthreads)
    dt = time.perf_counter() - t0
    assert int(r["status"].max()) == 0
    return n / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    co, cm, ph = cpu_port()
    rate, _ = cpu_solve_rate(cm, ph.names, 256, SEED, cores)                       # calibration
    sample = int(max(256, min(args.n_per_gpu, rate * 8.0)))                       # ~8 s of wall per step on all cores
    for w in range(args.warmup):
        cpu_solve_rate(cm, ph.names, max(64, sample // 8), SEED + 1 + w, cores)
    tot_t, tot_n = 0.0, 0
    for s in range(args.steps):
        _, dt = cpu_solve_rate(cm, ph.names, sample, SEED + 100 + s, cores)
        tot_t += dt
        tot_n += sample
    val = tot_n / tot_t
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "H2/O2 (superminimal) ConstantVDomain IDT ensemble, configs[1]", "conditions_per_step": sample,
                       "full_workload_conditions": args.n_per_gpu, "t_end": T_END, "rtol": RTOL, "atol": ATOL},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} of {args.n_per_gpu} conditions per step (same seeded sweep); " + CPU_BUILD_NOTE},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import rmsb200 as rb
    from rmsb200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        args.gpus = world
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = L.lib()
    if lib.rmsb_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: librmsb200 has no CPU fallback")

    d = rb.readinput(FIXTURE)
    ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase", device=local)
    n = args.n_per_gpu
    Ny, Ns, Nr = ig.Ns + 2, ig.Ns, ig.Nr
    y0, V = conditions(ig.names, n, SEED + 1000 * rank)
    te = np.full(n, T_END)
    h = ig.handle
    KIND = L.DOMAIN_CONSTANT_V

    # ---- device-resident buffers (value): inputs already in HBM when the timed region starts
    t_y0, t_V, t_te = (torch.from_numpy(a).to(dev) for a in (y0, V, te))
    t_yend = torch.empty((n, Ny), dtype=torch.float64, device=dev)
    t_idt = torch.empty(n, dtype=torch.float64, device=dev)
    t_status = torch.empty(n, dtype=torch.int32, device=dev)
    t_stats = torch.empty((n, 10), dtype=torch.int32, device=dev)          # rmsb_stats = 40 bytes
    lib.rmsb_solve_work_bytes.restype = C.c_int64
    wbytes = int(lib.rmsb_solve_work_bytes(h, KIND, n))                       # work-queue counter + the hottest-first work order
    t_work = torch.zeros((wbytes + 7) // 8, dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    opts = L.SolveOpts(C.sizeof(L.SolveOpts), 500000, 5, -1, 0.0, args.warps_per_block, args.flags)
    g_idt = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    g_st = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    def step_dev():
        L.check(lib.rmsb_solve_batch_dev(h, KIND, n, t_y0.data_ptr(), t_V.data_ptr(), None, 0, None, 0, t_te.data_ptr(), RTOL, ATOL,
                                         0, None, None, t_yend.data_ptr(), t_idt.data_ptr(), t_status.data_ptr(), t_stats.data_ptr(),
                                         C.byref(opts), t_work.data_ptr(), sp))
        if world > 1:   # the one exchange step of the path: collect ignition delays / status on rank 0 over NVLink
            dist.gather(t_idt, g_idt, dst=0)
            dist.gather(t_status, g_st, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        launches0 = lib.rmsb_kernel_launch_count()
        evs = []
        t_wall0 = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1.0)                               # L2 flush between timed iterations (outside the events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        wall = time.perf_counter() - t_wall0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = lib.rmsb_kernel_launch_count() - launches0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, launches

    sampler = ClockSampler(local) if rank == 0 else None
    ms, wall, launches = timed(step_dev, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    value = world * n * args.steps / (ms * 1e-3)
    stats = t_stats.cpu().numpy()
    st_ok = int((t_status != 0).sum().item()) == 0
    mean = {k: float(stats[:, i].mean()) for i, k in enumerate(["n_steps", "n_rhs", "n_jac", "n_lu", "n_newton_fail", "n_err_fail"])}

    # ---- e2e: the public host-buffer API (pinned host memory in, results back on the host), copies inside the timed region
    p_y0 = torch.from_numpy(y0).pin_memory(); p_V = torch.from_numpy(V).pin_memory(); p_te = torch.from_numpy(te).pin_memory()
    p_yend = torch.empty((n, Ny), dtype=torch.float64).pin_memory(); p_idt = torch.empty(n, dtype=torch.float64).pin_memory()
    p_status = torch.empty(n, dtype=torch.int32).pin_memory()

    def step_host():
        L.check(lib.rmsb_solve_batch(h, KIND, n, p_y0.data_ptr(), p_V.data_ptr(), None, 0, None, 0, p_te.data_ptr(), RTOL, ATOL,
                                     0, None, None, p_yend.data_ptr(), p_idt.data_ptr(), p_status.data_ptr(), None, C.byref(opts)))

    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_val = world * n * e2e_steps / t_e2e
    # the same call with the pageable NumPy arrays a Julia / NumPy caller would pass (rank 0 view; one warm call, one timed)
    g_yend, g_idt, g_status = np.empty((n, Ny)), np.empty(n), np.empty(n, dtype=np.int32)

    def step_pageable():
        L.check(lib.rmsb_solve_batch(h, KIND, n, y0.ctypes.data, V.ctypes.data, None, 0, None, 0, te.ctypes.data, RTOL, ATOL,
                                     0, None, None, g_yend.ctypes.data, g_idt.ctypes.data, g_status.ctypes.data, None, C.byref(opts)))
    step_pageable()
    t0 = time.perf_counter()
    step_pageable()
    e2e_pageable = n / (time.perf_counter() - t0)
    h2d = 8 * (n * Ny + n + n)
    d2h = 8 * (n * Ny + n) + 4 * n

    # ---- secondary metric: dydt + Jacobian evals/sec (device-resident), with the HBM roofline of SURVEY 8(d)
    ne = 1 << 20
    reps = (ne + n - 1) // n
    t_ye = t_y0.repeat(reps, 1)[:ne].contiguous(); t_Ve = t_V.repeat(reps, 1)[:ne].contiguous()
    t_f = torch.empty((ne, Ny), dtype=torch.float64, device=dev)
    t_J = torch.empty((ne, Ny * Ny), dtype=torch.float64, device=dev)

    def ev(fn, k=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(k):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream)
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / k * 1e-3

    generated = bool(lib.rmsb_mech_specialized(h, KIND))      # compiles the mechanism's own evaluation kernels now (csrc/rmsb_spec.cu)
    t_dydt = ev(lambda: L.check(lib.rmsb_dydt_batch_dev(h, KIND, ne, t_ye.data_ptr(), t_Ve.data_ptr(), None, 0, None, 0, t_f.data_ptr(), sp)))
    t_jac = ev(lambda: L.check(lib.rmsb_jac_batch_dev(h, KIND, ne, t_ye.data_ptr(), t_Ve.data_ptr(), None, 0, None, 0, t_J.data_ptr(), sp)))
    hbm_peak, peak_src = measured_peaks()
    fp64_peak = lib.rmsb_fp64_peak_tflops(local, 1 << 16)
    b_dydt, b_jac = 8 * (2 * Ny + 1), 8 * (Ny + 1 + Ny * Ny)

    # ---- roofline of the dominant kernel (solve_kernel): algorithmic HBM bytes per solve, and the FP64 view that explains it
    bytes_per_solve = 8 * (Ny + 1 + 1) + 8 * (Ny + 1) + 4 + 40          # in: y0, V, t_end; out: yend, idt, status, stats
    t_kernel = ms * 1e-3 / args.steps                                     # one launch per step
    ach_gbs = bytes_per_solve * n / t_kernel / 1e9
    # algorithmic flops per solve: n_rhs*F_rhs + n_jac*F_jac + n_lu*(2/3 Ny^3) + n_newton*(2 Ny^2)
    F_rhs, F_jac = 1.9e3, 4.0e3 + 3 * 1.9e3
    flops = mean["n_rhs"] * F_rhs + mean["n_jac"] * F_jac + mean["n_lu"] * (2.0 / 3.0 * Ny ** 3) + mean["n_rhs"] * 2 * Ny * Ny
    ach_tf = flops * n / t_kernel / 1e12

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cores = os.cpu_count() or 1
            co, cm, ph = cpu_port()
            rate, _ = cpu_solve_rate(cm, ph.names, 256, SEED, cores)
            sample = int(max(256, min(n, rate * 10.0)))
            val, dt = cpu_solve_rate(cm, ph.names, sample, SEED, cores)
            cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {sample} of the {n} conditions, {dt:.1f} s on {cores} threads; " + CPU_BUILD_NOTE}
            try:       # the strict (checker) build of the same source on a smaller sample, for the record
                _, cm_s, _ = cpu_port(fast=False)
                s_n = int(max(256, sample // 8))
                cpu["strict_build_value"], _ = cpu_solve_rate(cm_s, ph.names, s_n, SEED, cores)
            except Exception:
                pass
        except Exception as exc:   # the checker failing must not take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}

    # ---- strong scaling: the SAME 100k-condition set split over the ranks (BASELINE quotes 100k conditions in total).  Reactors are
    #      sorted by the work the warm-up solve measured for them (n_rhs, a pilot pass) and dealt round-robin in sorted chunks, so every
    #      rank gets the same mix; the limiter that shows is per-reactor serial latency (the tail), not a collective.
    strong = None
    if world > 1:
        n_tot = n
        y0g, Vg = conditions(ig.names, n_tot, SEED)                # identical on every rank
        # pilot: rank r measures the work of its contiguous block, all-gathered
        lo, hi = n_tot * rank // world, n_tot * (rank + 1) // world
        m = hi - lo
        ty, tv, tte = torch.from_numpy(y0g[lo:hi]).to(dev), torch.from_numpy(Vg[lo:hi]).to(dev), torch.full((m,), T_END, dtype=torch.float64, device=dev)
        L.check(lib.rmsb_solve_batch_dev(h, KIND, m, ty.data_ptr(), tv.data_ptr(), None, 0, None, 0, tte.data_ptr(), RTOL, ATOL, 0, None, None,
                                         t_yend.data_ptr(), t_idt.data_ptr(), t_status.data_ptr(), t_stats.data_ptr(), C.byref(opts), t_work.data_ptr(), sp))
        torch.cuda.synchronize()
        mx = max(n_tot * (r + 1) // world - n_tot * r // world for r in range(world))
        work_local = torch.zeros(mx, dtype=torch.int32, device=dev)
        work_local[:m] = t_stats[:m, 1]
        parts = [torch.empty(mx, dtype=torch.int32, device=dev) for _ in range(world)]
        dist.all_gather(parts, work_local)
        work_all = np.concatenate([parts[r][:n_tot * (r + 1) // world - n_tot * r // world].cpu().numpy() for r in range(world)])
        from rmsb200.ensemble import balanced_shards
        mine = balanced_shards(work_all, world, 64)[rank]           # heaviest first, round-robin of sorted chunks
        ms_by = {}
        # the sorted split keeps its cost order inside the library too (flags bit 6): neighbours in the work queue then behave alike,
        # which is worth more to the lockstep reactor pairs of a warp than the library's own hottest-first order
        opts_keep = L.SolveOpts(C.sizeof(L.SolveOpts), 500000, 5, -1, 0.0, args.warps_per_block, args.flags | 64)
        for name, sel in (("sorted_round_robin", mine), ("contiguous", np.arange(lo, hi))):
            o_ref = C.byref(opts_keep if name == "sorted_round_robin" else opts)
            ty, tv = torch.from_numpy(np.ascontiguousarray(y0g[sel])).to(dev), torch.from_numpy(np.ascontiguousarray(Vg[sel])).to(dev)
            mm = len(sel)
            tte = torch.full((mm,), T_END, dtype=torch.float64, device=dev)

            def step_strong():
                L.check(lib.rmsb_solve_batch_dev(h, KIND, mm, ty.data_ptr(), tv.data_ptr(), None, 0, None, 0, tte.data_ptr(), RTOL, ATOL, 0, None, None,
                                                 t_yend.data_ptr(), t_idt.data_ptr(), t_status.data_ptr(), t_stats.data_ptr(), o_ref,
                                                 t_work.data_ptr(), sp))
            k_strong = max(3, args.steps)
            ms_s, _, _ = timed(step_strong, k_strong, 2)
            ms_by[name] = (ms_s / k_strong, k_strong)
        strong = {"conditions_total": n_tot, "split": "stiffness sort (n_rhs of a pilot solve, heaviest first) + round-robin of 64-reactor chunks, order kept inside the "
                                                            "library (flags bit 6); contiguous_split_value = contiguous blocks in the library's own hottest-first order",
                  "value": n_tot / (ms_by["sorted_round_robin"][0] * 1e-3), "ms_per_step": ms_by["sorted_round_robin"][0],
                  "contiguous_split_value": n_tot / (ms_by["contiguous"][0] * 1e-3), "unit": UNIT, "timing": "CUDA events, max over ranks",
                  "limiter": "per-reactor serial latency: 12.5k reactors per GPU at N=8 are ~2.6 waves of the 4736 resident lane groups, "
                             "the last wave runs partly empty; no collective on the data path"}

    others = None
    if rank == 0 and world == 1 and not args.no_others:
        try:
            sys.path.insert(0, str(ROOT / "scripts"))
            import bench_configs
            others = bench_configs.run_all()
        except Exception as exc:
            others = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "H2/O2 (superminimal) ConstantVDomain IDT ensemble, configs[1]", "conditions_per_gpu": n,
                       "Ny": Ny, "n_species": Ns, "n_reactions": Nr, "t_end": T_END, "rtol": RTOL, "atol": ATOL,
                       "integrator": "variable-order NDF/BDF 1-5, one 16-lane group per reactor (two reactors per warp), mechanism tables in shared memory", "l2": "flushed between timed steps (256 MiB fill)",
                       "all_converged": st_ok, "mean_per_solve": mean, "wall_s_timed_region": wall},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "rmsb_solve_batch (host buffers, pinned)", "pageable_value": e2e_pageable,
                    "pageable_note": "same call with pageable NumPy arrays (what a Julia / NumPy caller passes), rank 0, one timed call"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "traffic": _ncu_traffic(n), "kernel": "small_solve_kernel<16>", "algorithmic_bytes_per_solve": bytes_per_solve,
                         "peak_source": peak_src,
                         "note": "state lives in registers/shared memory for the whole integration: the solve is FP64-pipe/latency bound, "
                                 "not HBM bound (SURVEY 8d); see roofline_fp64 and evals"},
            "roofline_fp64": {"achieved_tflops": ach_tf, "peak_tflops": fp64_peak, "frac": ach_tf / fp64_peak if fp64_peak > 0 else None,
                              "flops_per_solve": flops, "peak_source": "rmsb_fp64_peak_tflops (DFMA microbenchmark, this run)"},
            "roofline_issue": _issue_roofline(mean["n_steps"], n, t_kernel, clocks, local),
            "evals": {"n": ne,
                      "kernels": ("rmsb_spec_dydt / rmsb_spec_jac: generated for the mechanism, compiled with NVRTC at first use" if generated
                                  else "dydt_tpr_kernel / jac_tpr_kernel (generic; NVRTC unavailable or RMSB_NO_SPEC=1)"),
                      "dydt": {"evals_per_s": ne / t_dydt, "GBps": b_dydt * ne / t_dydt / 1e9, "frac_hbm": b_dydt * ne / t_dydt / 1e9 / hbm_peak,
                               "bytes_per_eval": b_dydt},
                      "jac": {"evals_per_s": ne / t_jac, "GBps": b_jac * ne / t_jac / 1e9, "frac_hbm": b_jac * ne / t_jac / 1e9 / hbm_peak,
                              "bytes_per_eval": b_jac}},
            "cpu_baseline": cpu,
            "others": others,
            "scaling_strong": strong,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n-per-gpu", type=int, default=100000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--warps-per-block", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the `others` block (the BASELINE configs besides the headline one)")
    ap.add_argument("--flags", type=int, default=0, help="integrator flags (rmsb_solve_opts.reserved); 0 = defaults")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    import __graft_entry__ as ge
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        ge.build_cuda()
        ge.build_oracle()
    if args.impl == "reference":
        run_reference(args)
    else:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            time.sleep(0.5 if int(os.environ.get("LOCAL_RANK", "0")) else 0.0)
        run_ours(args)


if __name__ == "__main__":
    main()
