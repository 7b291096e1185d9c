#!/usr/bin/env python
"""All BASELINE.json configs on ONE GPU: time, algorithmic GB/s (or TFLOP/s), fraction of the
measured roofline.  JSON lines on stdout.
    python benchmarks/configs.py [--only C2,C3] [--small]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import lbm_b200 as L  # noqa: E402

PEAK = 6551.7
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


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


def emit(name, ms, byts=None, flops=None, **kw):
    r = {"config": name, "ms": ms}
    if byts is not None:
        r["alg_GB"] = byts / 1e9
        r["GBs"] = byts / ms / 1e6
        r["frac_hbm_measured"] = r["GBs"] / PEAK
    if flops is not None:
        r["TFLOPs"] = flops / ms / 1e9
    r.update(kw)
    print(json.dumps(r), flush=True)


def vec(n, seed, dt=torch.float64):
    t = torch.empty(n, dtype=dt, device="cuda")
    L.fill_uniform(t, seed)
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="M,C1,C2,C3,GM,MM,BC,RD,SP,C4,C5")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--quick", action="store_true", help="1 warm-up + 1 timed launch per config (profiling runs)")
    a = ap.parse_args()
    only = a.only.split(",")
    if a.quick:
        global timeit
        _t = timeit
        timeit = lambda fn, reps=1, warm=1: _t(fn, 1, 1)  # noqa: E731
    sh = 3 if a.small else 0
    h = L.handle_for(0)

    if "M" in only:
        n = 1 << (27 - sh)
        for (l, u) in ((1, 1), (8, 8)):
            A = L.brand(n, n, l, u)
            x, y = vec(n, 9), torch.empty(n, dtype=torch.float64, device="cuda")
            emit(f"M gbmv f64 n=2^{27 - sh} l=u={l}", timeit(lambda: L.mul_(y, A, x)), 8 * n * (l + u + 3))
            emit(f"M gbmv^T f64 n=2^{27 - sh} l=u={l}", timeit(lambda: L.mul_(y, A.T, x)), 8 * n * (l + u + 3))
            del A, x, y
            torch.cuda.empty_cache()
        n = 1 << (25 - sh)
        A = L.brand(n, n, 128, 128)
        x, y = vec(n, 9), torch.empty(n, dtype=torch.float64, device="cuda")
        emit(f"M gbmv f64 n=2^{25 - sh} l=u=128 (n=2^27 needs >= 2 GPUs)", timeit(lambda: L.mul_(y, A, x), 5, 2), 8 * n * 259)
        emit(f"M gbmv^T f64 n=2^{25 - sh} l=u=128", timeit(lambda: L.mul_(y, A.T, x), 5, 2), 8 * n * 259)
        del A, x, y
        torch.cuda.empty_cache()

    if "C1" in only:
        n = 10_000
        A = L.brand(n, n, 1, 1)
        M = L.ApplyMatrix("*", A, A)
        x, y = vec(n, 9), torch.empty(n, dtype=torch.float64, device="cuda")
        Cm = M.materialize()
        emit("C1 materialize A*A n=10^4 (1,1)->(2,2)", timeit(lambda: L.gbmm(A, A, out=Cm), 50, 5), 8 * n * 11, note="launch-latency bound: report us")
        emit("C1 mul!(y, ApplyMatrix(*,A,A), x) n=10^4", timeit(lambda: M.mul_(y, x), 50, 5), 8 * n * 10, note="two gbmv launches + a temporary")

    if "C2" in only:
        n = 1 << (28 - sh)
        d, e, x = vec(n, 1), vec(n - 1, 2), vec(n, 4)
        y = torch.empty(n, dtype=torch.float64, device="cuda")
        for nm, A, arr in (("SymTridiagonal", L.SymTridiagonal(d, e), 4), ("Bidiagonal U", L.Bidiagonal(d, e, "U"), 4),
                           ("Bidiagonal L", L.Bidiagonal(d, e, "L"), 4)):
            emit(f"C2 {nm} mul! f64 n=2^{28 - sh}", timeit(lambda: L.mul_(y, A, x)), arr * 8 * n)
        e2 = vec(n - 1, 3)
        T = L.Tridiagonal(e, d, e2)
        emit(f"C2 Tridiagonal mul! f64 n=2^{28 - sh}", timeit(lambda: L.mul_(y, T, x)), 5 * 8 * n)
        h.set_tuning(tri_force_shuffle=1)
        emit(f"C2 SymTridiagonal (shuffle kernel) n=2^{28 - sh}", timeit(lambda: L.mul_(y, L.SymTridiagonal(d, e), x)), 4 * 8 * n)
        h.set_tuning(tri_force_shuffle=0)
        del d, e, e2, x, y
        torch.cuda.empty_cache()
        # mat-mat: 8 right-hand sides share one read of the diagonals
        n2, nrhs = 1 << (24 - sh), 8
        d, e, e2 = vec(n2, 1), vec(n2 - 1, 2), vec(n2 - 1, 3)
        X = torch.empty((nrhs, n2), dtype=torch.float64, device="cuda")
        L.fill_uniform(X, 9)
        Y = torch.empty_like(X)
        for nm, A, arr in (("SymTridiagonal", L.SymTridiagonal(d, e), 2), ("Tridiagonal", L.Tridiagonal(e, d, e2), 3),
                           ("Bidiagonal U", L.Bidiagonal(d, e, "U"), 2)):
            emit(f"C2 {nm} * dense f64 n=2^{24 - sh} nrhs={nrhs}", timeit(lambda: L.mul_(Y.t(), A, X.t())), 8 * n2 * (arr + 2 * nrhs))
        del d, e, e2, X, Y
        torch.cuda.empty_cache()

    if "C3" in only:
        g = 8192 >> sh
        hh = 1.0 / g
        D = L.banded_from_diagonals(g, {0: -2.0 / hh**2, 1: 1.0 / hh**2, -1: 1.0 / hh**2})
        Lap = L.KronSum(D, D)
        x = vec(g * g, 5)
        y = torch.empty_like(x)
        best = None
        for tr, tc in (() if a.quick else ((256, 16), (256, 32), (512, 16), (128, 32), (1024, 8), (256, 8), (512, 32))):
            h.set_tuning(kron_tr=tr, kron_tc=tc)
            ms = timeit(lambda: Lap.mul_(y, x))
            emit(f"C3 2-D Laplacian kronsum {g}x{g} tile {tr}x{tc}", ms, 16 * g * g)
            if best is None or ms < best[0]:
                best = (ms, tr, tc)
        h.set_tuning(kron_tr=0, kron_tc=0)
        emit(f"C3 2-D Laplacian kronsum {g}x{g} default", timeit(lambda: Lap.mul_(y, x)), 16 * g * g,
             best_tile=best[1:] if best else None)
        del x, y
        torch.cuda.empty_cache()

    if "GM" in only:
        # narrow banded x banded materialise at a size that fills the machine (north-star: >= 70 % of HBM roofline, l,u <= 8)
        n = 1 << (24 - sh)
        for dt, w in ((torch.float64, 8), (torch.float32, 4)):
            for l in (1, 2, 4, 8):
                A = L.brand(n, n, l, l, dtype=dt, seed=1)
                B = L.brand(n, n, l, l, dtype=dt, seed=2)
                Cm = L.BandedMatrix(torch.empty((n, 4 * l + 1), dtype=dt, device="cuda"), n, 2 * l, 2 * l)
                byts = w * n * ((2 * l + 1) * 2 + 4 * l + 1)
                flops = 2 * n * (2 * l + 1) ** 2
                nm = "f64" if w == 8 else "f32"
                emit(f"GM gbmm {nm} n=2^{24 - sh} ({l},{l})x({l},{l}) register kernel", timeit(lambda: L.gbmm(A, B, out=Cm)), byts, flops)
                if not a.quick:
                    h.set_tuning(gbmm_force=1)
                    emit(f"GM gbmm {nm} n=2^{24 - sh} ({l},{l})x({l},{l}) diagonal-wise kernel (gbmm_force=1)",
                         timeit(lambda: L.gbmm(A, B, out=Cm), 3, 1), byts, flops)
                    h.set_tuning(gbmm_force=0)
                del A, B, Cm
        torch.cuda.empty_cache()

    if "MM" in only:
        # banded x dense (north_star "mat-mat applies"): 8 right-hand sides share one pass over A
        n, nrhs = 1 << (24 - sh), 8
        for l in (1, 8):
            A = L.brand(n, n, l, l, seed=1)
            X = torch.empty((nrhs, n), dtype=torch.float64, device="cuda")
            L.fill_uniform(X, 9)
            Y = torch.empty_like(X)
            byts = 8 * n * ((2 * l + 1) + 2 * nrhs)
            ms = timeit(lambda: L.mul_(Y.t(), A, X.t()))
            emit(f"MM banded*dense f64 n=2^{24 - sh} l=u={l} nrhs={nrhs} (lbm_dgbmv_multi, one pass over A)", ms, byts)
            if not a.quick:
                for tile in (256, 512, 768, 1024, 1536):
                    h.set_tuning(gbmv_tile=tile)
                    emit(f"MM banded*dense f64 n=2^{24 - sh} l=u={l} nrhs={nrhs} tile={tile}", timeit(lambda: L.mul_(Y.t(), A, X.t())), byts)
                h.set_tuning(gbmv_tile=0)
            ms1 = timeit(lambda: [L.mul_(Y[c], A, X[c]) for c in range(nrhs)])
            emit(f"MM banded*dense f64 n=2^{24 - sh} l=u={l} nrhs={nrhs} (column by column, {nrhs} gbmv launches)", ms1, byts,
                 note="same algorithmic bytes credited; A re-read per column")
            del A, X, Y
        torch.cuda.empty_cache()

    if "RD" in only:
        # single-pass reductions on the structured kinds (8(f) rank 3)
        n = 1 << (28 - sh)
        S = L.SymTridiagonal(vec(n, 2), vec(n - 1, 3))
        x, y = vec(n, 4), vec(n, 5)
        emit(f"RD dot(x, SymTridiagonal, y) f64 n=2^{28 - sh}", timeit(lambda: S.dot(x, y)), 8 * 4 * n)
        emit(f"RD sum(SymTridiagonal; dims=1) f64 n=2^{28 - sh}", timeit(lambda: S.sum(1)), 8 * 3 * n)
        emit(f"RD sum(SymTridiagonal) f64 n=2^{28 - sh}", timeit(lambda: S.sum()), 8 * 2 * n)
        del S, x, y
        torch.cuda.empty_cache()

    if "SP" in only:
        # `+` among structured kinds (src/special.jl:75-222) and their BandedMatrix storage / products
        n = 1 << (27 - sh)
        T1 = L.Tridiagonal(vec(n - 1, 1), vec(n, 2), vec(n - 1, 3))
        T2 = L.Tridiagonal(vec(n - 1, 4), vec(n, 5), vec(n - 1, 6))
        Bu = L.Bidiagonal(vec(n, 7), vec(n - 1, 8), "U")
        emit(f"SP Tridiagonal + Tridiagonal f64 n=2^{27 - sh} (lbm_dtridiag_axpby, 3 diagonals in one launch)",
             timeit(lambda: T1 + T2), 8 * 9 * n)
        emit(f"SP Bidiagonal - Tridiagonal f64 n=2^{27 - sh}", timeit(lambda: Bu - T2), 8 * (2 + 2 + 2 + 2) * n)
        del T2, Bu
        torch.cuda.empty_cache()
        emit(f"SP BandedMatrix(Tridiagonal) f64 n=2^{27 - sh} (lbm_dtridiag_pack)", timeit(lambda: L.to_banded(T1)), 8 * 6 * n)
        n2 = 1 << (24 - sh)
        S = L.SymTridiagonal(vec(n2, 2), vec(n2 - 1, 3))
        Bl = L.Bidiagonal(vec(n2, 7), vec(n2 - 1, 8), "L")
        emit(f"SP SymTridiagonal * Bidiagonal -> BandedMatrix (2,1) f64 n=2^{24 - sh} (lbm_dtridiag_mm, one pass)",
             timeit(lambda: S @ Bl), 8 * n2 * (2 + 2 + 4))
        del T1, S, Bl
        torch.cuda.empty_cache()

    if "BC" in only:
        # BroadcastArray(+, A1, A2) * x, fused (lbm_dgbmv2) vs term by term (two gbmv, y read + written twice)
        n = 1 << (27 - sh)
        A1 = L.brand(n, n, 1, 1, seed=1)
        A2 = L.brand(n, n, 1, 1, seed=2)
        x, y = vec(n, 9), torch.empty(n, dtype=torch.float64, device="cuda")
        M = L.BroadcastMatrix("+", A1, A2)
        byts = 8 * n * (3 + 3 + 2)
        emit(f"BC (A1+A2)*x f64 n=2^{27 - sh} (1,1)+(1,1) fused lbm_dgbmv2", timeit(lambda: M.mul_(y, x)), byts)
        if not a.quick:
            for tile, stages, ctas in ((512, 2, 0), (768, 2, 0), (1024, 2, 0), (1536, 2, 0), (2048, 2, 0), (512, 3, 0), (1024, 3, 0), (1024, 2, 1), (2048, 2, 1), (1024, 4, 1)):
                h.set_tuning(gbmv_tile=tile, gbmv_stages=stages, gbmv_ctas_per_sm=ctas)
                emit(f"BC fused tile={tile} stages={stages} ctas={ctas or 'auto'}", timeit(lambda: M.mul_(y, x)), byts)
            h.set_tuning(gbmv_tile=0, gbmv_stages=0, gbmv_ctas_per_sm=0)
        emit(f"BC (A1+A2)*x f64 n=2^{27 - sh} (1,1)+(1,1) term by term (2 gbmv)",
             timeit(lambda: (L.mul_(y, A1, x), L.mul_(y, A2, x, 1.0, 1.0))), byts, note="same algorithmic bytes credited")
        del A1, A2, x, y
        torch.cuda.empty_cache()

    if "C5" in only:
        batch, n = 4096 >> sh, 4096
        fs = [torch.empty((batch, n, 9), dtype=torch.float32, device="cuda") for _ in range(3)]
        for i, f in enumerate(fs):
            L.fill_uniform(f, 1 + i)
        byts = batch * 4 * n * (9 + 9 + 9 + 25)
        flops = batch * 2 * n * (9 * 9 + 17 * 9)
        variants = ((0, "default = warp-autonomous TMA kernel, single-buffered, 8 warps/SM"),) if a.quick else (
            (0, "default = warp-autonomous TMA kernel, single-buffered, 8 warps/SM"),
            (5, "warp-autonomous TMA kernel, A/B double-buffered, 6 warps/SM"), (1, "split-column register kernel"),
            (2, "transposed-smem kernel, 204 regs / 2 CTAs"), (3, "transposed-smem kernel, 128 regs / 3 CTAs"))
        for v, nm in variants:
            h.set_tuning(chain_variant=v)
            emit(f"C5 fused chain A*B*C f32 batch={batch} n=4096 (4,4)^3 variant {v} ({nm})",
                 timeit(lambda: L.chain3_batched(*fs, [4, 4, 4], [4, 4, 4]), 10, 3), byts, flops)
        h.set_tuning(chain_variant=0)
        del fs
        torch.cuda.empty_cache()

    if "PEAK" in only:
        # FP64 denominator: cuBLAS DGEMM (library call, measurement only) + a plain device copy
        N = 8192
        a_ = torch.randn(N, N, dtype=torch.float64, device="cuda")
        b_ = torch.randn(N, N, dtype=torch.float64, device="cuda")
        c_ = torch.empty_like(a_)
        ms = timeit(lambda: torch.matmul(a_, b_, out=c_), 5, 2)
        emit("PEAK cuBLAS DGEMM 8192^3 (FP64 denominator)", ms, None, 2 * N**3)
        del a_, b_, c_
        src = torch.empty(1 << 29, dtype=torch.float64, device="cuda")
        dst = torch.empty_like(src)
        emit("PEAK torch copy 4 GiB f64 (HBM denominator)", timeit(lambda: dst.copy_(src)), 2 * src.numel() * 8)
        del src, dst
        torch.cuda.empty_cache()

    if "C4" in only:
        n = 1 << (20 - sh)
        A = L.brand(n, n, 128, 128, seed=1)
        B = L.brand(n, n, 128, 128, seed=2)
        Cm = L.BandedMatrix(torch.empty((n, 513), dtype=torch.float64, device="cuda"), n, 256, 256)
        flops = 2 * n * 257 * 257
        byts = 8 * n * (257 + 257 + 513)
        for force in ((0,) if a.quick else (1, 0)):
            h.set_tuning(gbmm_force=force)
            ms = timeit(lambda: L.gbmm(A, B, out=Cm), 2, 1)
            emit(f"C4 gbmm f64 n=2^{20 - sh} (128,128)^2 gbmm_force={force}", ms, byts, flops)
        h.set_tuning(gbmm_force=0)
        if not a.quick:
            for v, nm in ((1, "2-stage, 4 CTAs/SM"), (2, "3-stage, 3 CTAs/SM"), (3, "2-stage, 5 CTAs/SM"), (4, "column-persistent 2-stage, 4 CTAs/SM"),
                          (5, "column-persistent 3-stage, 3 CTAs/SM"), (6, "3-stage, 4 CTAs/SM"), (7, "4-stage, 3 CTAs/SM"),
                          (8, "3-stage, 4 CTAs/SM, rotated quadrants"), (9, "3-stage, 4 CTAs/SM, interleaved p-halves"), (10, "3-stage, 4 CTAs/SM, both"),
                          (11, "3-stage, 4 CTAs/SM, single launch over all tiles"), (12, "2-stage, 6 CTAs/SM"),
                          (20, "warp-specialised: producer warp + 8 consumer warps, 3-stage, 4 CTAs/SM"),
                          (21, "warp-specialised, 3-stage, 3 CTAs/SM"), (22, "warp-specialised, 4-stage, 3 CTAs/SM"),
                          (23, "warp-specialised, 2-stage, 4 CTAs/SM")):
                h.set_tuning(gbmm_variant=v)
                emit(f"C4 gbmm f64 n=2^{20 - sh} (128,128)^2 DMMA variant {v} ({nm})", timeit(lambda: L.gbmm(A, B, out=Cm), 2, 1), byts, flops)
            for tpc in (2, 3, 5, 10):
                h.set_tuning(gbmm_variant=22, gbmm_tile=tpc)
                emit(f"C4 gbmm f64 n=2^{20 - sh} (128,128)^2 DMMA variant 22 (warp-specialised, 4-stage, 3 CTAs/SM), {tpc} row tiles per CTA",
                     timeit(lambda: L.gbmm(A, B, out=Cm), 2, 1), byts, flops)
                h.set_tuning(gbmm_variant=20, gbmm_tile=tpc)
                emit(f"C4 gbmm f64 n=2^{20 - sh} (128,128)^2 DMMA variant 20 (warp-specialised, 3-stage, 4 CTAs/SM), {tpc} row tiles per CTA",
                     timeit(lambda: L.gbmm(A, B, out=Cm), 2, 1), byts, flops)
            for code, nm in ((205, "32 ns poll back-off"), (305, "128 ns poll back-off"), (5, "default")):
                h.set_tuning(gbmm_variant=0, gbmm_tile=code)
                emit(f"C4 gbmm f64 n=2^{20 - sh} (128,128)^2 DMMA default kernel, {nm}", timeit(lambda: L.gbmm(A, B, out=Cm), 3, 1), byts, flops)
            h.set_tuning(gbmm_variant=0, gbmm_tile=0)


if __name__ == "__main__":
    main()
