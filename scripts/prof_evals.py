"""The dydt / Jacobian evaluation kernels alone on the bench's 2^20 C2 states (for `ncu --set full -k regex:tpr_kernel`).
Usage: python scripts/prof_evals.py [reps]"""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "reactionmechanismsimulator.jl_b200"):
    sys.path.insert(0, str(p))
import bench  # noqa: E402
import rmsb200 as rb  # noqa: E402
from rmsb200 import _lib as L  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3   # ncu: -s 2 -c 2 captures the third dydt launch and the first Jacobian launch
lib = L.lib()
d = rb.readinput(bench.FIXTURE)
ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
ne = 1 << 20
y0, V = bench.conditions(ig.names, ne, bench.SEED)
Ny = ig.Ns + 2
dev = torch.device("cuda", 0)
ty, tv = torch.from_numpy(y0).to(dev), torch.from_numpy(V).to(dev)
tf = torch.empty((ne, Ny), dtype=torch.float64, device=dev)
tJ = torch.empty((ne, Ny * Ny), dtype=torch.float64, device=dev)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
h = ig.handle
print('specialised kernels:', ig.specialized(L.DOMAIN_CONSTANT_V), lib.rmsb_last_error().decode() if not ig.specialized(L.DOMAIN_CONSTANT_V) else '')
for name, fn in (("dydt", lambda: lib.rmsb_dydt_batch_dev(h, L.DOMAIN_CONSTANT_V, ne, ty.data_ptr(), tv.data_ptr(), None, 0, None, 0, tf.data_ptr(), sp)),
                 ("jac", lambda: lib.rmsb_jac_batch_dev(h, L.DOMAIN_CONSTANT_V, ne, ty.data_ptr(), tv.data_ptr(), None, 0, None, 0, tJ.data_ptr(), sp))):
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.check(fn()); e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name}: {ms:.3f} ms for {ne} states -> {ne / ms * 1e3:.3e} evals/s")
