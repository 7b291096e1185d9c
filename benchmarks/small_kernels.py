#!/usr/bin/env python
"""Per-rank-sized kernels on ONE GPU (what one rank of an 8-GPU job runs): back-to-back launch time vs the ideal, and the
tuning knobs that matter at that size.  JSON lines.  python benchmarks/small_kernels.py [--only K,G,T,S]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import lbm_b200 as L  # noqa: E402

PEAK = 6551.7


def timeit(fn, reps=200, warm=20):
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


def emit(name, ms, byts, **kw):
    r = {"config": name, "us": ms * 1e3, "GBs": byts / ms / 1e6, "frac": byts / ms / 1e6 / PEAK, "ideal_us": byts / PEAK / 1e3}
    r.update(kw)
    print(json.dumps(r), flush=True)


def main():
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else ["K", "G", "T", "S"]
    h = L.handle_for(0)
    if "K" in only:
        g = 8192
        D = L.banded_from_diagonals(g, {0: -2.0, 1: 1.0, -1: 1.0})
        for n1 in (1024, 2048, 8192):
            A = L.banded_from_diagonals(n1, {0: -2.0, 1: 1.0, -1: 1.0})
            Lap = L.KronSum(A, D)
            x = torch.rand(n1 * g, dtype=torch.float64, device="cuda")
            y = torch.empty_like(x)
            emit(f"kronsum n1={n1} n2={g} default", timeit(lambda: Lap.mul_(y, x)), 16 * n1 * g)
            if n1 == 1024:
                for tc in (4, 8, 16, 32):
                    for st in (2, 3, 4):
                        for cps in (0, 2, 3, 4):
                            h.set_tuning(kron_tc=tc, kron_stages=st, kron_ctas_per_sm=cps)
                            emit(f"kronsum n1={n1} tc={tc} stages={st} cps={cps}", timeit(lambda: Lap.mul_(y, x), 100, 10), 16 * n1 * g)
                for tr in (512, 1024):
                    h.set_tuning(kron_tr=tr, kron_tc=8, kron_stages=2, kron_ctas_per_sm=0)
                    emit(f"kronsum n1={n1} tr={tr} tc=8", timeit(lambda: Lap.mul_(y, x), 100, 10), 16 * n1 * g)
                h.set_tuning(kron_tr=0, kron_tc=0, kron_stages=0, kron_ctas_per_sm=0)
    if "G" in only:
        for (l, log2n) in ((1, 24), (8, 24), (1, 27)):
            n = 1 << log2n
            A = L.brand(n, n, l, l)
            x = torch.rand(n, dtype=torch.float64, device="cuda")
            y = torch.empty_like(x)
            byts = 8 * n * (2 * l + 3)
            emit(f"gbmv n=2^{log2n} l=u={l} default", timeit(lambda: L.mul_(y, A, x)), byts)
            if log2n == 24:
                for tile in (256, 512, 1024, 2048):
                    for st in (2, 3, 4):
                        for cps in (0, 2, 4, 6):
                            h.set_tuning(gbmv_tile=tile, gbmv_stages=st, gbmv_ctas_per_sm=cps)
                            emit(f"gbmv n=2^{log2n} l=u={l} tile={tile} stages={st} cps={cps}", timeit(lambda: L.mul_(y, A, x), 100, 10), byts)
                h.set_tuning(gbmv_tile=0, gbmv_stages=0, gbmv_ctas_per_sm=0)
            del A, x, y
            torch.cuda.empty_cache()
    if "B" in only:
        # the per-rank slab of a row shard is a RECTANGULAR banded matrix with bandwidths (l - left, u + left): (1,1) on rank 0,
        # (0,2) on the others -- same 3 stored diagonals, different alignment of the slab / x window inside a tile
        n = 1 << 25
        for (kl, ku) in ((1, 1), (0, 2), (2, 0)):
            data = torch.rand((n + 2, 3), dtype=torch.float64, device="cuda")
            A = L.BandedMatrix(data, n, kl, ku)          # n x (n+2)
            x = torch.rand(n + 2, dtype=torch.float64, device="cuda")
            y = torch.empty(n, dtype=torch.float64, device="cuda")
            emit(f"gbmv rectangular n=2^25 x (n+2), bandwidths ({kl},{ku})", timeit(lambda: L.mul_(y, A, x)), 8 * n * 5)
            if (kl, ku) != (2, 0):
                for tile in (512, 768, 1024, 1280, 1536, 2048):
                    for st in (2, 3):
                        for cps in (0, 2, 4):
                            h.set_tuning(gbmv_tile=tile, gbmv_stages=st, gbmv_ctas_per_sm=cps)
                            emit(f"gbmv rectangular ({kl},{ku}) tile={tile} stages={st} cps={cps}", timeit(lambda: L.mul_(y, A, x), 50, 5), 8 * n * 5)
                h.set_tuning(gbmv_tile=0, gbmv_stages=0, gbmv_ctas_per_sm=0)
                for skew in (2, 4, 6, 16):
                    h.set_tuning(gbmv_skew=skew)
                    emit(f"gbmv rectangular ({kl},{ku}) skew={skew}", timeit(lambda: L.mul_(y, A, x), 50, 5), 8 * n * 5)
                h.set_tuning(gbmv_skew=0)
            del data, A, x, y
            torch.cuda.empty_cache()
    if "P" in only:
        # programmatic dependent launch off (0, default) / on (1) at the sizes one rank of an 8-GPU job runs
        g = 8192
        D = L.banded_from_diagonals(g, {0: -2.0, 1: 1.0, -1: 1.0})
        A1 = L.banded_from_diagonals(1024, {0: -2.0, 1: 1.0, -1: 1.0})
        Lap = L.KronSum(A1, D)
        xk = torch.rand(1024 * g, dtype=torch.float64, device="cuda")
        yk = torch.empty_like(xk)
        n = 1 << 24
        G1, G8 = L.brand(n, n, 1, 1), L.brand(n, n, 8, 8)
        xg = torch.rand(n, dtype=torch.float64, device="cuda")
        yg = torch.empty_like(xg)
        nt = 1 << 25
        S = L.SymTridiagonal(torch.rand(nt, dtype=torch.float64, device="cuda"), torch.rand(nt - 1, dtype=torch.float64, device="cuda"))
        xt = torch.rand(nt, dtype=torch.float64, device="cuda")
        yt = torch.empty_like(xt)
        for rep in range(2):
            for pdl in (0, 1):
                h.set_tuning(pdl=pdl)
                tag = "pdl on" if pdl == 1 else "pdl off"
                emit(f"kronsum 1024x8192 {tag}", timeit(lambda: Lap.mul_(yk, xk)), 16 * 1024 * g)
                emit(f"gbmv n=2^24 l=u=1 {tag}", timeit(lambda: L.mul_(yg, G1, xg)), 8 * n * 5)
                emit(f"gbmv n=2^24 l=u=8 {tag}", timeit(lambda: L.mul_(yg, G8, xg)), 8 * n * 19)
                emit(f"symtridiag n=2^25 {tag}", timeit(lambda: L.mul_(yt, S, xt)), 32 * nt)
                # ping-pong: each apply reads what the previous one wrote (the dependency PDL must respect)
                def pp():
                    L.mul_(yg, G1, xg)
                    L.mul_(xg, G1, yg)
                emit(f"gbmv n=2^24 l=u=1 ping-pong pair {tag}", timeit(pp, 100, 10) / 2, 8 * n * 5)
        h.set_tuning(pdl=0)
    if "T" in only:
        for log2n in (25, 28):
            n = 1 << log2n
            d = torch.rand(n, dtype=torch.float64, device="cuda")
            e = torch.rand(n - 1, dtype=torch.float64, device="cuda")
            x = torch.rand(n, dtype=torch.float64, device="cuda")
            y = torch.empty_like(x)
            S = L.SymTridiagonal(d, e)
            emit(f"symtridiag mv n=2^{log2n}", timeit(lambda: L.mul_(y, S, x)), 32 * n)
            del d, e, x, y, S
            torch.cuda.empty_cache()
    if "S" in only:
        for dt, w in ((torch.float64, 8), (torch.float32, 4)):
            for log2n in (20, 24, 27):
                n = 1 << log2n
                d = torch.rand(n, dtype=dt, device="cuda") + 4
                e = torch.rand(n - 1, dtype=dt, device="cuda")
                b = torch.rand(n, dtype=dt, device="cuda")
                x = torch.empty_like(b)
                S = L.SymTridiagonal(d, e)
                T = L.Tridiagonal(e, d, e.clone())
                Bi = L.Bidiagonal(d, e, "U")
                emit(f"solve SymTridiagonal {dt} n=2^{log2n} (5n = one read of ev twice,dv,b + write x)", timeit(lambda: S.solve(b, out=x, check=False), 20, 3), 5 * w * n)
                emit(f"solve Tridiagonal {dt} n=2^{log2n}", timeit(lambda: T.solve(b, out=x, check=False), 20, 3), 5 * w * n)
                emit(f"solve Bidiagonal U {dt} n=2^{log2n}", timeit(lambda: Bi.solve(b, out=x, check=False), 20, 3), 4 * w * n)
                del d, e, b, x, S, T, Bi
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
