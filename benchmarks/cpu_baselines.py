#!/usr/bin/env python
"""CPU baselines on the box's host cores for every BASELINE config (reduced sizes, stated): OpenBLAS ?gbmv via scipy
(the routine the reference's gbmv! ccalls), 'reference-style' gbmm = n column-wise dgbmv calls (upstream gbmm!'s
algorithm), and the repo's OpenMP oracle for the paths with no BLAS equivalent.  JSON lines."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from scipy.linalg import blas  # noqa: E402

from oracle import oracle as O  # noqa: E402


def best(fn, reps=3):
    fn()
    t = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t = min(t, time.perf_counter() - t0)
    return t


def emit(name, s, byts=None, flops=None, **kw):
    r = {"config": name, "ms": s * 1e3, "threads": O.num_threads()}
    if byts is not None:
        r["GBs"] = byts / s / 1e9
    if flops is not None:
        r["GFLOPs"] = flops / s / 1e9
    r.update(kw)
    print(json.dumps(r), flush=True)


def main():
    n = 1 << 24
    for (l, u) in ((1, 1), (8, 8)):
        A, x, y = O.brand(n, n, l, u, 1), O.fill_uniform(9, n), np.zeros(n)
        emit(f"M dgbmv OpenBLAS n=2^24 l=u={l}", best(lambda: blas.dgbmv(n, n, l, u, 1.0, A, x, beta=0.0, y=y, overwrite_y=1)), 8 * n * (l + u + 3))
        emit(f"M gbmv OpenMP oracle n=2^24 l=u={l}", best(lambda: O.gbmv("N", n, n, l, u, 1.0, A, x)), 8 * n * (l + u + 3))
    n = 1 << 20
    A, x, y = O.brand(n, n, 128, 128, 1), O.fill_uniform(9, n), np.zeros(n)
    emit("M dgbmv OpenBLAS n=2^20 l=u=128", best(lambda: blas.dgbmv(n, n, 128, 128, 1.0, A, x, beta=0.0, y=y, overwrite_y=1)), 8 * n * 259)
    emit("M gbmv OpenMP oracle n=2^20 l=u=128", best(lambda: O.gbmv("N", n, n, 128, 128, 1.0, A, x)), 8 * n * 259)
    n = 1 << 24
    d, e, x = O.fill_uniform(1, n), O.fill_uniform(2, n - 1), O.fill_uniform(4, n)
    emit("C2 SymTridiagonal OpenMP oracle n=2^24", best(lambda: O.tridiag_mv(n, e, d, e, x)), 4 * 8 * n)
    emit("C2 Bidiagonal U OpenMP oracle n=2^24", best(lambda: O.tridiag_mv(n, None, d, e, x)), 4 * 8 * n)
    emit("C2 numpy 3-slice SymTridiagonal n=2^24", best(lambda: d * x + np.concatenate(([0.0], e * x[:-1])) + np.concatenate((e * x[1:], [0.0]))), 4 * 8 * n)
    g = 2048
    h = 1.0 / g
    D, _, _ = O.banded_from_diagonals(g, {0: -2.0 / h**2, 1: 1.0 / h**2, -1: 1.0 / h**2})
    x = O.fill_uniform(5, g * g)
    emit("C3 kronsum OpenMP oracle 2048^2", best(lambda: O.kronsum2d_mv(g, g, D, 1, 1, D, 1, 1, x)), 16 * g * g)
    n = 10_000
    A = O.brand(n, n, 1, 1, 1)
    emit("C1 gbmm OpenMP oracle n=10^4 (1,1)^2", best(lambda: O.gbmm(n, n, n, A, 1, 1, A, 1, 1)), 8 * n * 11)

    def ref_style_gbmm(A, la, ua, B, lb, ub, n):
        """upstream gbmm!: one gbmv per column of B against the sub-band of A it touches"""
        C = np.zeros((la + lb + ua + ub + 1, n), order="F")
        for j in range(n):
            p0, p1 = max(0, j - ub), min(n, j + lb + 1)
            i0, i1 = max(0, p0 - ua), min(n, p1 + la)
            # rows i0:i1 of A[:, p0:p1] as a banded block: column slice with shifted bandwidths
            sh = i0 - p0
            blk = A[:, p0:p1]
            yj = blas.dgbmv(i1 - i0, p1 - p0, la - sh, ua + sh, 1.0, blk, B[ub + p0 - j: ub + p1 - j, j]) if (la - sh >= 0 and i1 - i0 >= la + ua + 1) else \
                O.gbmv("N", i1 - i0, p1 - p0, la - sh, ua + sh, 1.0, np.asfortranarray(blk), np.ascontiguousarray(B[ub + p0 - j: ub + p1 - j, j]), lda=blk.shape[0])
            C[ua + ub + i0 - j: ua + ub + i1 - j, j] = yj
        return C
    t = best(lambda: ref_style_gbmm(A, 1, 1, A, 1, 1, n), 1)
    emit("C1 gbmm reference-style (n column dgbmv calls) n=10^4", t, 8 * n * 11)
    n = 1 << 13
    A, B = O.brand(n, n, 128, 128, 1), O.brand(n, n, 128, 128, 2)
    emit("C4 gbmm OpenMP oracle n=2^13 l=u=128", best(lambda: O.gbmm(n, n, n, A, 128, 128, B, 128, 128), 1), 8 * n * 1027, 2 * n * 257 * 257)
    fs = [O.fill_uniform(1 + i, 64 * 4096 * 9, np.float32).reshape(64, 4096, 9) for i in range(3)]
    emit("C5 chain OpenMP oracle batch=64 n=4096 f32", best(lambda: O.gbmm_chain3_batched(4096, [4] * 3, [4] * 3, *fs), 1), 64 * 4 * 4096 * 52, 64 * 2 * 4096 * 234)


if __name__ == "__main__":
    main()
