"""Generated (csrc/rmsb_spec.cu) vs generic evaluation kernels on synthetic all-Arrhenius gas mechanisms of growing size.
Usage: python scripts/gpu_spec_sizes.py  (run once plain and once with RMSB_NO_SPEC=1)"""
import ctypes as C
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import rmsb200 as rb  # noqa: E402
import synth  # noqa: E402
from rmsb200 import _lib as L  # noqa: E402

lib = L.lib()
dev = torch.device("cuda", 0)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
n = 1 << 18
for ns, rps in ((18, 2.0), (20, 3.0), (24, 3.0), (27, 4.0), (30, 4.0)):
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "m.rms")
        synth.write_synthetic_mechanism(f, ns=ns, rxn_per_species=rps, seed=11)
        d = rb.readinput(f)
    ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
    rng = np.random.default_rng(1)
    Ny = ig.Ns + 2
    y = np.empty((n, Ny))
    y[:, :ig.Ns] = rng.uniform(0.01, 1.0, (n, ig.Ns))
    y[:, ig.Ns] = rng.uniform(900.0, 1600.0, n)
    y[:, ig.Ns + 1] = rng.uniform(1e5, 2e6, n)
    V = np.full(n, 1.0)
    ty, tv = torch.from_numpy(y).to(dev), torch.from_numpy(V).to(dev)
    tf = torch.empty((n, Ny), dtype=torch.float64, device=dev)
    tJ = torch.empty((n, Ny * Ny), dtype=torch.float64, device=dev)
    h = ig.handle
    t0 = time.time()
    spec = ig.specialized(L.DOMAIN_CONSTANT_V)
    tc = time.time() - t0
    out = [f"Ns {ig.Ns:3d} Nr {ig.Nr:4d} generated={spec} (compile {tc:5.1f} s)"]
    for name, fn in (("dydt", lambda: lib.rmsb_dydt_batch_dev(h, L.DOMAIN_CONSTANT_V, n, ty.data_ptr(), tv.data_ptr(), None, 0, None, 0, tf.data_ptr(), sp)),
                     ("jac", lambda: lib.rmsb_jac_batch_dev(h, L.DOMAIN_CONSTANT_V, n, ty.data_ptr(), tv.data_ptr(), None, 0, None, 0, tJ.data_ptr(), sp))):
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); L.check(fn()); e1.record()
            torch.cuda.synchronize()
        out.append(f"{name} {n / e0.elapsed_time(e1) * 1e3:.3e}/s")
    out.append(f"checksum {float(tf.nan_to_num().abs().sum()):.6e} {float(tJ.nan_to_num().abs().sum()):.6e}")
    print("  ".join(out), flush=True)
