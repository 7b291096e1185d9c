#!/usr/bin/env python
"""lbm_sgbmm with l=u=128: the f64 tensor route (widen, DMMA, round) vs the diagonal-wise kernel.  JSON lines."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L

n = 1 << int(os.environ.get("LOG2N", "18"))
A = L.brand(n, n, 128, 128, dtype=torch.float32, seed=1)
B = L.brand(n, n, 128, 128, dtype=torch.float32, seed=2)
C = L.BandedMatrix(torch.empty((n, 513), dtype=torch.float32, device="cuda"), n, 256, 256)
h = L.handle_for(0)
flops = 2 * n * 257 * 257
for force, name in ((0, "f64 tensor route (default)"), (1, "diagonal-wise f32 kernel")):
    h.set_tuning(gbmm_force=force)
    for _ in range(2):
        L.gbmm(A, B, out=C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        L.gbmm(A, B, out=C)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"config": f"sgbmm f32 n=2^{n.bit_length() - 1} l=u=128, {name}", "ms": round(ms, 3), "TFLOPs": round(flops / ms / 1e9, 2)}), flush=True)
h.set_tuning(gbmm_force=0)
