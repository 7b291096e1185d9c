"""Pins the C oracle's gbmv against OpenBLAS ?gbmv (scipy.linalg.blas) -- the BLAS routine the
reference's `gbmv!` ccalls -- and gbmm / tridiag against dense numpy on small sizes."""
import numpy as np
import pytest
from scipy.linalg import blas

from tests._util import tol_rows

SHAPES = [  # (m, n, kl, ku)
    (1, 1, 0, 0), (5, 5, 1, 1), (7, 7, 0, 2), (9, 9, 3, 0), (64, 64, 8, 8), (33, 40, 2, 5), (40, 33, 5, 2),
    (3, 3, 4, 4), (2, 9, 1, 1), (9, 2, 1, 1), (300, 300, 128, 128), (10, 4, 0, 0), (1000, 1000, 1, 1),
]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("trans", ["N", "T"])
def test_gbmv_vs_openblas(oracle, dt, trans):
    f = blas.dgbmv if dt == np.float64 else blas.sgbmv
    for s, (m, n, kl, ku) in enumerate(SHAPES):
        A = oracle.brand(m, n, kl, ku, seed=10 + s, dtype=dt)
        nin, nout = (n, m) if trans == "N" else (m, n)
        x = oracle.fill_uniform(99 + s, nin, dt)
        y0 = oracle.fill_uniform(77 + s, nout, dt)
        for alpha, beta in [(1.0, 0.0), (2.0, 1.0), (-0.5, 0.25)]:
            if m >= kl + ku + 1:
                want = f(m, n, kl, ku, alpha, A, x, beta=beta, y=y0.copy(), trans=1 if trans == "T" else 0)
            else:  # scipy's f2py wrapper refuses m < kl+ku+1 (BLAS itself does not): dense check
                D = oracle.band_to_dense(A, m, kl, ku).astype(np.float64)
                want = alpha * ((D if trans == "N" else D.T) @ x.astype(np.float64)) + beta * y0
            got = oracle.gbmv(trans, m, n, kl, ku, alpha, A, x, beta, y0)
            tol = tol_rows(dt, kl + ku + 1, 1.0, 1.0, alpha, beta, 1.0)
            assert np.max(np.abs(got - want)) <= tol, (m, n, kl, ku, alpha, beta)


def test_gbmv_ignores_out_of_matrix_slots_and_beta0_nan(oracle):
    m = n = 12
    kl, ku = 2, 3
    A = oracle.brand(m, n, kl, ku, seed=5)
    clean = oracle.gbmv("N", m, n, kl, ku, 1.0, A, np.ones(n))
    for j in range(n):       # poison every out-of-matrix slot
        for r in range(kl + ku + 1):
            i = j + r - ku
            if i < 0 or i >= m:
                A[r, j] = np.nan
    y0 = np.full(m, np.nan)
    got = oracle.gbmv("N", m, n, kl, ku, 1.0, A, np.ones(n), 0.0, y0)
    assert np.array_equal(got, clean)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_gbmm_vs_dense(oracle, dt):
    rng = np.random.default_rng(0)
    cases = [(6, 6, 6, 1, 1, 1, 1), (10, 8, 9, 2, 0, 1, 3), (5, 5, 5, 4, 4, 4, 4), (12, 12, 12, 0, 0, 2, 1),
             (1, 1, 1, 0, 0, 0, 0), (20, 17, 23, 3, 2, 1, 4)]
    for (m, n, k, kla, kua, klb, kub) in cases:
        A = oracle.brand(m, k, kla, kua, seed=int(rng.integers(1 << 30)), dtype=dt)
        B = oracle.brand(k, n, klb, kub, seed=int(rng.integers(1 << 30)), dtype=dt)
        Cb = oracle.gbmm(m, n, k, A, kla, kua, B, klb, kub)
        want = oracle.band_to_dense(A, m, kla, kua).astype(np.float64) @ oracle.band_to_dense(B, k, klb, kub).astype(np.float64)
        got = oracle.band_to_dense(Cb, m, kla + klb, kua + kub)
        tol = tol_rows(dt, min(kla + kua + 1, klb + kub + 1), 1.0, 1.0)
        assert np.max(np.abs(got - want)) <= tol
        # the product really is confined to the (l1+l2, u1+u2) band
        assert np.max(np.abs(want - oracle.band_to_dense(oracle.dense_to_band(want, kla + klb, kua + kub), m, kla + klb, kua + kub))) == 0


def test_apply_chain_readme_case(oracle):
    """README.md:13-26: A = brand(10,10,1,1); ApplyMatrix(*, A, A) is pentadiagonal (2,2)."""
    A = oracle.brand(10, 10, 1, 1, seed=1)
    C = oracle.gbmm(10, 10, 10, A, 1, 1, A, 1, 1)
    assert C.shape == (5, 10)
    dense = oracle.band_to_dense(A, 10, 1, 1)
    assert np.allclose(oracle.band_to_dense(C, 10, 2, 2), dense @ dense, rtol=0, atol=1e-15)
    x = oracle.fill_uniform(9, 10)
    y = oracle.gbmv("N", 10, 10, 1, 1, 1.0, A, oracle.gbmv("N", 10, 10, 1, 1, 1.0, A, x))
    assert np.allclose(y, dense @ (dense @ x), rtol=0, atol=1e-15)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_tridiag_vs_dense(oracle, dt):
    for n in (1, 2, 3, 17, 256):
        d = oracle.fill_uniform(1, n, dt)
        dl = oracle.fill_uniform(2, max(n - 1, 0), dt)
        du = oracle.fill_uniform(3, max(n - 1, 0), dt)
        X = oracle.fill_uniform(4, n * 2, dt).reshape(n, 2, order="F")
        dense = np.diag(d.astype(np.float64))
        if n > 1:
            dense += np.diag(dl.astype(np.float64), -1) + np.diag(du.astype(np.float64), 1)
        for (a, b) in [(dl, du), (None, du), (dl, None), (dl, dl)]:
            D = np.diag(d.astype(np.float64))
            if n > 1:
                if a is not None:
                    D = D + np.diag(a.astype(np.float64), -1)
                if b is not None:
                    D = D + np.diag(b.astype(np.float64), 1)
            got = oracle.tridiag_mv(n, a if n > 1 else None, d, b if n > 1 else None, X)
            assert np.max(np.abs(got - D @ X.astype(np.float64))) <= tol_rows(dt, 3, 1.0, 1.0)
        # generalized dot (test/test_tridiag.jl:294-298): dot(x, A, y) == dot(A'x, y)
        if n > 1 and dt == np.float64:
            x = X[:, 0]
            y = X[:, 1]
            assert abs(oracle.tridiag_dot(x, dl, d, du, y) - (dense.T @ x) @ y) <= 1e-12


def test_chain3_batched_vs_dense(oracle):
    dt = np.float32
    batch, n = 3, 40
    kl, ku = [4, 4, 4], [4, 4, 4]
    A = oracle.fill_uniform(1, batch * n * 9, dt).reshape(batch, n, 9)
    B = oracle.fill_uniform(2, batch * n * 9, dt).reshape(batch, n, 9)
    Cm = oracle.fill_uniform(3, batch * n * 9, dt).reshape(batch, n, 9)
    D = oracle.gbmm_chain3_batched(n, kl, ku, A, B, Cm)
    assert D.shape == (batch, n, 25)
    for b in range(batch):
        dA, dB, dC = (oracle.band_to_dense(np.asfortranarray(M[b].T), n, 4, 4).astype(np.float64) for M in (A, B, Cm))
        got = oracle.band_to_dense(np.asfortranarray(D[b].T), n, 12, 12)
        assert np.max(np.abs(got - dA @ dB @ dC)) <= 4 * np.finfo(dt).eps * (9 * 9 + 17 * 9)
