#!/usr/bin/env python
"""C4 (l=u=128, n=2^20) with a list of (gbmm_variant, gbmm_tile) tuning pairs: python benchmarks/c4_quick.py 0:5 0:205 0:305"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L

n = 1 << 20
A = L.brand(n, n, 128, 128, seed=1)
B = L.brand(n, n, 128, 128, seed=2)
Cm = L.BandedMatrix(torch.empty((n, 513), dtype=torch.float64, device="cuda"), n, 256, 256)
h = L.handle_for(0)
flops = 2 * n * 257 * 257
pairs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(0, 0)]
for rep in range(2):
    for v, t in pairs:
        h.set_tuning(gbmm_variant=v, gbmm_tile=t)
        for _ in range(2):
            L.gbmm(A, B, out=Cm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            L.gbmm(A, B, out=Cm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({"variant": v, "tile": t, "ms": round(ms, 3), "TFLOPs": round(flops / ms / 1e9, 2)}), flush=True)
