#!/usr/bin/env python
"""Tolerance report: for each operation of the path, max observed |gpu - oracle| over the
bound 4*eps*(terms)*||A||*||x|| (max-norms; stage bounds add for chains).  Markdown table on stdout."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import lbm_b200 as L  # noqa: E402
from oracle import oracle as O  # noqa: E402

rows = []


def band(m, n, l, u, seed, dt=np.float64):
    a = O.brand(m, n, l, u, seed=seed, dtype=dt)
    return a, L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(a.T)).cuda(), m, l, u)


def rec(name, err, bound):
    rows.append((name, err, bound, err / bound))


def main():
    for dt in (np.float64, np.float32):
        eps = np.finfo(dt).eps
        nm = "f64" if dt == np.float64 else "f32"
        for (n, l, u) in ((1 << 20, 1, 1), (1 << 20, 8, 8), (1 << 17, 128, 128)):
            A_np, A = band(n, n, l, u, 1, dt)
            x = O.fill_uniform(9, n, dt)
            for tr in ("N", "T"):
                y = (A if tr == "N" else A.T) @ torch.from_numpy(x).cuda()
                want = O.gbmv(tr, n, n, l, u, 1.0, A_np, x)
                rec(f"gbmv {tr} {nm} n=2^{n.bit_length() - 1} l=u={l}", float(np.max(np.abs(y.cpu().numpy() - want))), 4 * eps * (l + u + 1))
        n = 1 << 20
        d, e = O.fill_uniform(1, n, dt), O.fill_uniform(2, n - 1, dt)
        x = O.fill_uniform(4, n, dt)
        t = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
        for name, Aop, (a, b) in (("SymTridiagonal", L.SymTridiagonal(t(d), t(e)), (e, e)), ("Bidiagonal U", L.Bidiagonal(t(d), t(e), "U"), (None, e)),
                                  ("Bidiagonal L", L.Bidiagonal(t(d), t(e), "L"), (e, None))):
            y = Aop @ t(x)
            rec(f"{name} {nm} n=2^20", float(np.max(np.abs(y.cpu().numpy() - O.tridiag_mv(n, a, d, b, x)))), 4 * eps * 3)
    eps = np.finfo(np.float64).eps
    # gbmm narrow (config 1) and wide DMMA (config 4 at reduced n)
    for (n, l, u) in ((10_000, 1, 1), (4096, 128, 128)):
        A_np, A = band(n, n, l, u, 1)
        B_np, B = band(n, n, l, u, 2)
        C = A @ B
        want = O.gbmm(n, n, n, A_np, l, u, B_np, l, u)
        got = np.asfortranarray(C.data.cpu().numpy().T)
        err = float(np.max(np.abs(O.band_to_dense(got, n, 2 * l, 2 * u) - O.band_to_dense(want, n, 2 * l, 2 * u))))
        rec(f"gbmm f64 n={n} (l=u={l})^2" + (" [DMMA]" if l >= 16 else ""), err, 4 * eps * (l + u + 1))
    # lazy chain apply (config 1): two stages
    n = 10_000
    A_np, A = band(n, n, 1, 1, 1)
    x = O.fill_uniform(9, n)
    y = L.ApplyMatrix("*", A, A) @ torch.from_numpy(x).cuda()
    want = O.gbmv("N", n, n, 1, 1, 1.0, A_np, O.gbmv("N", n, n, 1, 1, 1.0, A_np, x))
    rec("mul!(y, ApplyMatrix(*,A,A), x) n=10^4", float(np.max(np.abs(y.cpu().numpy() - want))), 2 * 4 * eps * 3 * 3.0)
    # batched fused chain (config 5 at reduced batch)
    e32 = np.finfo(np.float32).eps
    fs = [O.fill_uniform(1 + i, 8 * 4096 * 9, np.float32).reshape(8, 4096, 9) for i in range(3)]
    want = O.gbmm_chain3_batched(4096, [4] * 3, [4] * 3, *fs)
    got = L.chain3_batched(*[torch.from_numpy(f).cuda() for f in fs], [4] * 3, [4] * 3).cpu().numpy()
    err = max(float(np.max(np.abs(O.band_to_dense(np.asfortranarray(got[b].T), 4096, 12, 12) - O.band_to_dense(np.asfortranarray(want[b].T), 4096, 12, 12)))) for b in range(8))
    rec("fused chain A*B*C f32 n=4096 (4,4)^3", err, 4 * e32 * 9 * 9 + 4 * e32 * 9 * 9.0)
    # 2-D Laplacian
    g = 1024
    h = 1.0 / g
    D = L.banded_from_diagonals(g, {0: -2.0 / h**2, 1: 1.0 / h**2, -1: 1.0 / h**2})
    x = O.fill_uniform(5, g * g)
    D_np = np.asfortranarray(D.data.cpu().numpy().T)
    y = L.KronSum(D, D) @ torch.from_numpy(x).cuda()
    rec("kron(D,I)+kron(I,D) f64 1024^2", float(np.max(np.abs(y.cpu().numpy() - O.kronsum2d_mv(g, g, D_np, 1, 1, D_np, 1, 1, x)))), 4 * eps * 6 * 2 * g * g)
    print("| operation | max observed error | bound 4 eps (terms) ‖A‖ ‖x‖ | observed / bound |")
    print("|---|---|---|---|")
    for name, err, bound, ratio in rows:
        print(f"| {name} | {err:.3e} | {bound:.3e} | {ratio:.3f} |")
    assert all(r[3] <= 1.0 for r in rows)


if __name__ == "__main__":
    main()
