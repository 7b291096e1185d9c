#!/usr/bin/env python
"""Kronecker applies beyond the tridiagonal Laplacian (VERDICT r1 item 8 gate): KronSum / BlockKron with (2,2) and (4,4)
factors at 8192^2, f64 and f32, and the three mode products of a 512^3 kron3d.  Algorithmic bytes: KronSum and kron2d read x
and write y once (2N words); kron3d makes three sweeps (6N words).  CUDA-event timed, JSON lines."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L
from lbm_b200.banded import _ft, _ptr

PEAK = 6551.7


def timeit(fn, reps=20, warm=3):
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
    print(json.dumps({"config": name, "ms": round(ms, 4), "GBs": round(byts / ms / 1e6, 1), "frac": round(byts / ms / 1e6 / PEAK, 3)}), flush=True)


g = 8192
for dt, w in ((torch.float64, 8), (torch.float32, 4)):
    x = torch.rand(g * g, dtype=dt, device="cuda")
    y = torch.empty_like(x)
    for bw in ((1, 1), (2, 2), (4, 4)):
        A = L.brand(g, g, bw[0], bw[1], dtype=dt, seed=3)
        B = L.brand(g, g, bw[0], bw[1], dtype=dt, seed=4)
        S = L.KronSum(A, B)
        emit(f"KronSum {dt} 8192^2 factors {bw}", timeit(lambda: S.mul_(y, x)), 2 * w * g * g)
        K = L.BlockKron(A, B)
        emit(f"BlockKron(A,B)*x {dt} 8192^2 factors {bw}", timeit(lambda: K.mul_(y, x)), 2 * w * g * g)
    del x, y
n = 512
for dt, w, nm in ((torch.float64, 8, "lbm_dkron3d_mv"), (torch.float32, 4, "lbm_skron3d_mv")):
    x = torch.rand(n ** 3, dtype=dt, device="cuda")
    y = torch.empty_like(x)
    h = L._lib.handle_of(x)
    fn = getattr(L.lib(), nm)
    ft = _ft(dt)
    for bw in ((1, 1), (2, 2)):
        A, B, C = (L.brand(n, n, bw[0], bw[1], dtype=dt, seed=5 + i) for i in range(3))

        def run():
            h.check(fn(h.ptr, n, A.l, A.u, _ptr(A.data), A.lda, B.l, B.u, _ptr(B.data), B.lda, C.l, C.u, _ptr(C.data), C.lda,
                       ft(1.0), _ptr(x), ft(0.0), _ptr(y)), "kron3d")
        emit(f"kron3d {dt} 512^3 factors {bw}: three mode products (6N words)", timeit(run, 10, 2), 6 * w * n ** 3)
