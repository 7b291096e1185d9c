#!/usr/bin/env python
"""One warm-up + one launch of a single config, for `ncu --set full` captures:
    python benchmarks/prof_one.py C4 [log2n]   |   C5   |   C3"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L

which = sys.argv[1] if len(sys.argv) > 1 else "C4"
if which == "C4":
    n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 18)
    A = L.brand(n, n, 128, 128, seed=1)
    B = L.brand(n, n, 128, 128, seed=2)
    Cm = L.BandedMatrix(torch.empty((n, 513), dtype=torch.float64, device="cuda"), n, 256, 256)
    for _ in range(2):
        L.gbmm(A, B, out=Cm)
elif which == "C5":
    batch, n = 4096, 4096
    fs = [torch.empty((batch, n, 9), dtype=torch.float32, device="cuda") for _ in range(3)]
    for i, f in enumerate(fs):
        L.fill_uniform(f, 1 + i)
    for _ in range(2):
        D = L.chain3_batched(*fs, [4, 4, 4], [4, 4, 4])
torch.cuda.synchronize()
print("ok")
