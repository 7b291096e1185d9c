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
    L.handle_for(0).set_tuning(gbmm_variant=int(os.environ.get("LBM_GBMM_VARIANT", "0")))
    for _ in range(2):
        L.gbmm(A, B, out=Cm)
elif which == "C5":
    batch, n = 4096, 4096
    fs = [torch.empty((batch, n, 9), dtype=torch.float32, device="cuda") for _ in range(3)]
    for i, f in enumerate(fs):
        L.fill_uniform(f, 1 + i)
    for _ in range(2):
        D = L.chain3_batched(*fs, [4, 4, 4], [4, 4, 4])
elif which == "C3":
    g = 8192
    hh = 1.0 / g
    D = L.banded_from_diagonals(g, {0: -2.0 / hh**2, 1: 1.0 / hh**2, -1: 1.0 / hh**2})
    Lap = L.KronSum.from_broadcast(L.BlockKron(D, L.Eye(g)), L.BlockKron(L.Eye(g), D))
    x = torch.empty(g * g, dtype=torch.float64, device="cuda")
    L.fill_uniform(x, 5)
    y = torch.empty_like(x)
    for _ in range(2):
        Lap.mul_(y, x)
elif which == "C2":
    n = 1 << 28
    dv = torch.empty(n, dtype=torch.float64, device="cuda")
    ev = torch.empty(n - 1, dtype=torch.float64, device="cuda")
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    for i, t in enumerate((dv, ev, x)):
        L.fill_uniform(t, 2 + i)
    y = torch.empty_like(x)
    S = L.SymTridiagonal(dv, ev)
    B = L.Bidiagonal(dv, ev, "U")
    for _ in range(2):
        S.mul_(y, x)
        B.mul_(y, x)
    S.dot(x, y)
elif which == "M":
    n = 1 << 27
    for l in (1, 8):
        A = L.brand(n, n, l, l, seed=1)
        x = torch.empty(n, dtype=torch.float64, device="cuda")
        L.fill_uniform(x, 9)
        y = torch.empty_like(x)
        for _ in range(2):
            L.mul_(y, A, x)
        del A
elif which == "GM":
    n = 1 << 24
    for l in (8, 4, 1):
        A = L.brand(n, n, l, l, seed=1)
        B = L.brand(n, n, l, l, seed=2)
        Cm = L.BandedMatrix(torch.empty((n, 4 * l + 1), dtype=torch.float64, device="cuda"), n, 2 * l, 2 * l)
        for _ in range(2):
            L.gbmm(A, B, out=Cm)
        del A, B, Cm
elif which == "SP":
    n = 1 << 26
    def v(k, s):
        t = torch.empty(k, dtype=torch.float64, device="cuda")
        L.fill_uniform(t, s)
        return t
    T1 = L.Tridiagonal(v(n - 1, 1), v(n, 2), v(n - 1, 3))
    T2 = L.Tridiagonal(v(n - 1, 4), v(n, 5), v(n - 1, 6))
    for _ in range(2):
        R = T1 + T2
        P = T1 @ T2
elif which == "MM1":
    n, nrhs = 1 << 24, 8
    A = L.brand(n, n, 1, 1, seed=1)
    X = torch.empty((nrhs, n), dtype=torch.float64, device="cuda")
    L.fill_uniform(X, 9)
    Y = torch.empty_like(X)
    d = torch.empty(n, dtype=torch.float64, device="cuda"); L.fill_uniform(d, 3)
    e = torch.empty(n - 1, dtype=torch.float64, device="cuda"); L.fill_uniform(e, 4)
    S = L.SymTridiagonal(d, e)
    for _ in range(2):
        L.mul_(Y.t(), A, X.t())
        L.mul_(Y.t(), S, X.t())
elif which == "MM":
    n, nrhs = 1 << 24, 8
    A = L.brand(n, n, 8, 8, seed=1)
    X = torch.empty((nrhs, n), dtype=torch.float64, device="cuda")
    L.fill_uniform(X, 9)
    Y = torch.empty_like(X)
    A2 = L.brand(n, n, 1, 1, seed=2)
    A3 = L.brand(n, n, 1, 1, seed=3)
    M = L.BroadcastMatrix("+", A2, A3)
    for _ in range(2):
        L.mul_(Y.t(), A, X.t())
        M.mul_(Y[0], X[0])
elif which == "RT":
    # banded x banded without an exact instantiation: the run-time-count variant of the register kernel
    n = 1 << 24
    for (la, ua, lb, ub) in ((3, 3, 1, 1), (5, 5, 8, 8)):
        A = L.brand(n, n, la, ua, seed=1)
        B = L.brand(n, n, lb, ub, seed=2)
        Cm = L.BandedMatrix(torch.empty((n, la + ua + lb + ub + 1), dtype=torch.float64, device="cuda"), n, la + lb, ua + ub)
        for _ in range(2):
            L.gbmm(A, B, out=Cm)
        del A, B, Cm
elif which == "SV":
    n = 1 << 26
    d = torch.rand(n, dtype=torch.float64, device="cuda") + 4
    e = torch.rand(n - 1, dtype=torch.float64, device="cuda")
    b = torch.rand(n, dtype=torch.float64, device="cuda")
    x = torch.empty_like(b)
    S = L.SymTridiagonal(d, e)
    for _ in range(2):
        S.solve(b, out=x, check=False)
elif which == "C3S":
    # the 1024-column slab one rank of eight holds, plain and with the fused-exchange instantiation (no neighbours)
    g, n1 = 8192, 1024
    from lbm_b200.shard import RowShard, ShardedKronSum
    D = L.banded_from_diagonals(g, {0: -2.0, 1: 1.0, -1: 1.0})
    A = L.banded_from_diagonals(n1, {0: -2.0, 1: 1.0, -1: 1.0})
    p = RowShard.plan(n1, 1, 1, 0, 1)
    sh = ShardedKronSum(n1, g, 1, 1, A.data[p.c0:p.c1].contiguous(), D, 0, 1)
    xb = sh.new_buffer(torch.float64, torch.device("cuda", 0))
    xb.uniform_()
    y = torch.empty((p.m_loc, g), dtype=torch.float64, device="cuda")
    h = L.handle_for(0)
    for fuse in (0, 3):
        h.set_tuning(exchange_fuse=fuse)
        for _ in range(2):
            sh.mul_(y, xb)
    h.set_tuning(exchange_fuse=0)
torch.cuda.synchronize()
print("ok")
