#!/usr/bin/env python
"""Generality sweep: gbmv (N and T) and gbmm over band widths that are NOT BASELINE configs, to expose performance
cliffs between kernel paths.  JSON lines.   python benchmarks/sweep_bands.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L

PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6551.7


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def emit(name, ms, byts):
    print(json.dumps({"config": name, "ms": round(ms, 4), "GBs": round(byts / ms / 1e6, 1), "frac_hbm_measured": round(byts / ms / 1e6 / PEAK, 3)}), flush=True)


h = L.handle_for(0)
for (l, u) in ((0, 0), (2, 2), (4, 4), (0, 3), (5, 1), (16, 16), (24, 8), (32, 32), (64, 64)):
    n = 1 << (26 if l + u <= 16 else 24)
    A = L.brand(n, n, l, u)
    x = torch.empty(n, dtype=torch.float64, device="cuda"); L.fill_uniform(x, 9)
    y = torch.empty_like(x)
    byts = 8 * n * (l + u + 3)
    emit(f"gbmv N f64 n=2^{n.bit_length() - 1} (l,u)=({l},{u})", timeit(lambda: L.mul_(y, A, x)), byts)
    emit(f"gbmv T f64 n=2^{n.bit_length() - 1} (l,u)=({l},{u})", timeit(lambda: L.mul_(y, A.T, x)), byts)
    del A, x, y
    torch.cuda.empty_cache()
n = 1 << 24
for (la, ua, lb, ub) in ((3, 2, 1, 4), (0, 1, 1, 0), (8, 8, 1, 1), (2, 2, 8, 8), (5, 5, 5, 5), (1, 0, 0, 1)):
    A = L.brand(n, n, la, ua, seed=1)
    B = L.brand(n, n, lb, ub, seed=2)
    C = L.BandedMatrix(torch.empty((n, la + lb + ua + ub + 1), dtype=torch.float64, device="cuda"), n, la + lb, ua + ub)
    byts = 8 * n * ((la + ua + 1) + (lb + ub + 1) + (la + lb + ua + ub + 1))
    emit(f"gbmm f64 n=2^24 ({la},{ua})x({lb},{ub})", timeit(lambda: L.gbmm(A, B, out=C)), byts)
    del A, B, C
    torch.cuda.empty_cache()
# C1-sized materialise: diagonal-wise kernel (default below 256*SMs columns) vs the register kernel forced
for n in (10_000, 20_000, 37_000):
    A = L.brand(n, n, 1, 1)
    C = L.BandedMatrix(torch.empty((n, 5), dtype=torch.float64, device="cuda"), n, 2, 2)
    for force in (0, 3):
        h.set_tuning(gbmm_force=force)
        emit(f"gbmm f64 n={n} (1,1)^2 gbmm_force={force}", timeit(lambda: L.gbmm(A, A, out=C), 100, 10), 8 * n * 11)
    h.set_tuning(gbmm_force=0)
