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
    # narrow gbmm at a size that takes the register kernel (band arrays compared directly; n too large for dense)
    for dt in (np.float64, np.float32):
        e_ = np.finfo(dt).eps
        for l in (1, 8):
            n = 1 << 17
            A_np, A = band(n, n, l, l, 1, dt)
            B_np, B = band(n, n, l, l, 2, dt)
            got = np.asfortranarray((A @ B).data.cpu().numpy().T)
            want = O.gbmm(n, n, n, A_np, l, l, B_np, l, l)
            r = np.arange(4 * l + 1)[:, None]
            i = np.arange(n)[None, :] + r - 2 * l
            inside = (i >= 0) & (i < n)
            rec(f"gbmm {'f64' if dt == np.float64 else 'f32'} n=2^17 (l=u={l})^2 [register kernel]",
                float(np.max(np.abs(np.where(inside, got - want, 0)))), 4 * e_ * (2 * l + 1))
    # structured x structured product and mat-mat with 8 right-hand sides
    n = 1 << 17
    d, e1, e2 = O.fill_uniform(1, n), O.fill_uniform(2, n - 1), O.fill_uniform(3, n - 1)
    t = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    P = L.Tridiagonal(t(e1), t(d), t(e2)) @ L.SymTridiagonal(t(d), t(e2))
    gotP = np.asfortranarray(P.data.cpu().numpy().T)
    TA = np.zeros((3, n)); TA[1] = d; TA[0, 1:] = e2; TA[2, :-1] = e1          # (1,1) band storage of Tridiagonal(e1,d,e2)
    TS = np.zeros((3, n)); TS[1] = d; TS[0, 1:] = e2; TS[2, :-1] = e2
    wantP = O.gbmm(n, n, n, np.asfortranarray(TA), 1, 1, np.asfortranarray(TS), 1, 1)
    r = np.arange(5)[:, None]; i = np.arange(n)[None, :] + r - 2
    rec("Tridiagonal * SymTridiagonal f64 n=2^17 [lbm_dtridiag_mm]", float(np.max(np.abs(np.where((i >= 0) & (i < n), gotP - wantP, 0)))), 4 * eps * 3)
    X = O.fill_uniform(9, 8 * n).reshape(8, n)
    Y = torch.empty((8, n), dtype=torch.float64, device="cuda")
    L.mul_(Y.t(), L.SymTridiagonal(t(d), t(e1)), t(X).t())
    rec("SymTridiagonal * dense (8 columns) f64 n=2^17", float(max(np.max(np.abs(Y[c].cpu().numpy() - O.tridiag_mv(n, e1, d, e1, X[c]))) for c in range(8))), 4 * eps * 3)
    A_np, A = band(n, n, 1, 1, 1)
    L.mul_(Y.t(), A, t(X).t())
    rec("BandedMatrix (1,1) * dense (8 columns) f64 n=2^17", float(max(np.max(np.abs(Y[c].cpu().numpy() - O.gbmv("N", n, n, 1, 1, 1.0, A_np, X[c]))) for c in range(8))), 4 * eps * 3)
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
