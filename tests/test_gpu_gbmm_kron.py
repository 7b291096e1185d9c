"""GPU parity: banded x banded materialize, lazy chains (ApplyMatrix), batched fused chains and
the never-materialised block-Kronecker applies, vs the oracle."""
import numpy as np
import pytest
import torch

from tests._util import tol_rows

pytestmark = pytest.mark.gpu
NP = {torch.float64: np.float64, torch.float32: np.float32}


def _band(L, oracle, m, n, l, u, seed, dt):
    A_np = oracle.brand(m, n, l, u, seed=seed, dtype=NP[dt])
    return A_np, L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(A_np.T)).cuda(), m, l, u)


def _cmp_band(oracle, got_bm, want_np, m, l, u, tol):
    got = np.asfortranarray(got_bm.data.cpu().numpy().T)
    dg = oracle.band_to_dense(got, m, l, u)
    dw = oracle.band_to_dense(want_np, m, l, u)
    err = np.max(np.abs(dg - dw)) if dg.size else 0.0
    assert err <= tol, (err, tol)
    # out-of-matrix slots are written as zeros (beta == 0)
    n = got.shape[1]
    for j in list(range(min(n, u + 1))) + list(range(max(0, n - l - 1), n)):
        for r in range(l + u + 1):
            i = j + r - u
            if i < 0 or i >= m:
                assert got[r, j] == 0.0


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_gbmm_shapes(L, oracle, dt):
    cases = [(6, 6, 6, 1, 1, 1, 1), (10, 8, 9, 2, 0, 1, 3), (5, 5, 5, 4, 4, 4, 4), (12, 12, 12, 0, 0, 2, 1),
             (1, 1, 1, 0, 0, 0, 0), (200, 170, 230, 3, 2, 1, 4), (10000, 10000, 10000, 1, 1, 1, 1),
             (3000, 3000, 3000, 8, 8, 8, 8), (2000, 2100, 1900, 16, 3, 5, 20), (1500, 1500, 1500, 40, 40, 40, 40)]
    for s, (m, n, k, kla, kua, klb, kub) in enumerate(cases):
        A_np, A = _band(L, oracle, m, k, kla, kua, 11 + s, dt)
        B_np, B = _band(L, oracle, k, n, klb, kub, 12 + s, dt)
        want = oracle.gbmm(m, n, k, A_np, kla, kua, B_np, klb, kub)
        C = A @ B
        assert C.bandwidths() == (kla + klb, kua + kub) and C.shape == (m, n)
        tol = tol_rows(NP[dt], min(kla + kua + 1, klb + kub + 1), 1.0, 1.0)
        _cmp_band(oracle, C, want, m, kla + klb, kua + kub, tol)


def test_gbmm_dmma_tensor_path(L, oracle):
    """Wide bands go to the FP64 DMMA tile kernel; gbmm_force=2 forces it on narrow ones too.
    Same inputs through the diagonal-wise kernel (gbmm_force=1) must agree with the oracle as well."""
    h = L.handle_for(0)
    cases = [  # (m, n, k, kla, kua, klb, kub, force)
        (1000, 1000, 1000, 128, 128, 128, 128, 0), (700, 650, 720, 40, 35, 33, 50, 0), (300, 280, 310, 3, 2, 1, 4, 2),
        (64, 64, 64, 1, 1, 1, 1, 2), (129, 67, 200, 16, 0, 0, 16, 2), (500, 500, 500, 64, 10, 10, 64, 2), (5, 5, 5, 4, 4, 4, 4, 2),
        # raw-staged interior tiles with asymmetric bands (odd lda, even ku: aligned) and an unaligned variant (odd ku)
        (900, 850, 920, 40, 36, 34, 50, 0), (900, 900, 900, 41, 35, 35, 49, 0), (1500, 1500, 1500, 200, 56, 16, 100, 0),
        # sizes with an interior range for the warp-autonomous kernel (m, n multiples of 32, even ku): symmetric wide, asymmetric
        (4096, 4096, 4096, 128, 128, 128, 128, 0), (2048, 2048, 2048, 40, 36, 34, 50, 0), (3072, 3072, 3072, 200, 56, 16, 100, 0),
    ]
    for s, (m, n, k, kla, kua, klb, kub, force) in enumerate(cases):
        A_np, A = _band(L, oracle, m, k, kla, kua, 31 + s, torch.float64)
        B_np, B = _band(L, oracle, k, n, klb, kub, 32 + s, torch.float64)
        want = oracle.gbmm(m, n, k, A_np, kla, kua, B_np, klb, kub)
        tol = tol_rows(np.float64, min(kla + kua + 1, klb + kub + 1), 1.0, 1.0)
        for f, variant in ((force, 0), (force, 1), (force, 2), (force, 3), (force, 4), (force, 5), (force, 3), (force, 6), (force, 7), (force, 10), (force, 11), (force, 21), (force, 22), (force, 24), (force, 25), (force, 26), (force, 30), (force, 31), (force, 32), (1, 0)):
            h.set_tuning(gbmm_force=f, gbmm_variant=variant)
            try:
                before = h.launch_count()
                C = A @ B
                torch.cuda.synchronize()
                nl = h.launch_count() - before
            finally:
                h.set_tuning(gbmm_force=0, gbmm_variant=0)
            # zero-corners + DMMA tile kernel(s: interior + corner instantiations) vs the single diag kernel
            assert nl == 1 if f == 1 else nl in (2, 3)
            _cmp_band(oracle, C, want, m, kla + klb, kua + kub, tol)
    # alpha/beta and an output band wider than the product band, tensor path
    m = n = k = 400
    A_np, A = _band(L, oracle, m, k, 40, 33, 1, torch.float64)
    B_np, B = _band(L, oracle, k, n, 35, 36, 2, torch.float64)
    C0_np, C0 = _band(L, oracle, m, n, 80, 75, 3, torch.float64)
    want = oracle.gbmm(m, n, k, A_np, 40, 33, B_np, 35, 36, klc=80, kuc=75, alpha=-1.5, beta=0.25, Cin=C0_np)
    L.gbmm(A, B, alpha=-1.5, out=C0, beta=0.25)
    got = np.asfortranarray(C0.data.cpu().numpy().T)
    dg, dw = oracle.band_to_dense(got, m, 80, 75), oracle.band_to_dense(want, m, 80, 75)
    assert np.max(np.abs(dg - dw)) <= tol_rows(np.float64, 70, 1.0, 1.0, 1.5, 0.25, 1.0)


def test_gbmm_f32_wide_bands_through_the_f64_tensor_path(L, oracle):
    """lbm_sgbmm with both bands >= 33 diagonals: A, B (and C when beta != 0) are widened into scratch, the FP64 DMMA kernel
    runs, the band is rounded back.  Against the f32 oracle within the f32 bound; NaNs parked in the out-of-matrix slots;
    alpha / beta with an output band wider than the product band; the diagonal-wise kernel (gbmm_force = 1) agrees."""
    h = L.handle_for(0)
    dt, npdt = torch.float32, np.float32
    for s_, (m, n, k, kla, kua, klb, kub) in enumerate([(1000, 1000, 1000, 128, 128, 128, 128), (700, 650, 720, 40, 35, 33, 50),
                                                         (2048, 2048, 2048, 40, 36, 34, 50), (300, 400, 350, 64, 10, 10, 64)]):
        A_np = oracle.brand(m, k, kla, kua, seed=70 + s_, dtype=npdt)
        B_np = oracle.brand(k, n, klb, kub, seed=75 + s_, dtype=npdt)
        want = oracle.gbmm(m, n, k, A_np, kla, kua, B_np, klb, kub)

        def poison(M, rows, l, u):
            P = M.copy()
            r = np.arange(l + u + 1)[:, None]
            i = np.arange(P.shape[1])[None, :] + r - u
            P[(i < 0) | (i >= rows)] = np.nan
            return L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(P.T)).cuda(), rows, l, u)

        A, B = poison(A_np, m, kla, kua), poison(B_np, k, klb, kub)
        tol = tol_rows(npdt, min(kla + kua + 1, klb + kub + 1), 1.0, 1.0)
        for force in (0, 1):
            h.set_tuning(gbmm_force=force)
            try:
                before = h.launch_count()
                C = A @ B
                torch.cuda.synchronize()
                nl = h.launch_count() - before
            finally:
                h.set_tuning(gbmm_force=0)
            assert (nl >= 5) if force == 0 else (nl == 1), (force, nl)      # widen x2, zero corners, DMMA interior + corners, round
            assert C.data.dtype == torch.float32
            _cmp_band(oracle, C, want, m, kla + klb, kua + kub, tol)
    m = n = k = 400
    A_np, A = _band(L, oracle, m, k, 40, 33, 1, dt)
    B_np, B = _band(L, oracle, k, n, 35, 36, 2, dt)
    C0_np, C0 = _band(L, oracle, m, n, 80, 75, 3, dt)
    want = oracle.gbmm(m, n, k, A_np, 40, 33, B_np, 35, 36, klc=80, kuc=75, alpha=-1.5, beta=0.25, Cin=C0_np)
    L.gbmm(A, B, alpha=-1.5, out=C0, beta=0.25)
    got = np.asfortranarray(C0.data.cpu().numpy().T)
    dg, dw = oracle.band_to_dense(got, m, 80, 75), oracle.band_to_dense(want, m, 80, 75)
    assert np.max(np.abs(dg - dw)) <= tol_rows(npdt, 70, 1.0, 1.0, 1.5, 0.25, 1.0)


def test_gbmm_dmma_nan_poisoned_out_of_matrix_slots(L, oracle):
    """Interior tiles of the DMMA kernel stage RAW dense windows (out-of-band positions alias neighbouring
    columns' slots, including out-of-matrix ones) and mask in registers: NaNs parked in every out-of-matrix
    slot of A and B must not reach any in-matrix entry of C."""
    n, l = 1100, 128
    A_np, _ = _band(L, oracle, n, n, l, l, 5, torch.float64)
    B_np, _ = _band(L, oracle, n, n, l, l, 6, torch.float64)
    want = oracle.gbmm(n, n, n, A_np, l, l, B_np, l, l)
    Ap, Bp = A_np.copy(), B_np.copy()
    r = np.arange(2 * l + 1)[:, None]
    i = np.arange(n)[None, :] + r - l
    for M in (Ap, Bp):
        M[(i < 0) | (i >= n)] = np.nan
    A = L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(Ap.T)).cuda(), n, l, l)
    B = L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(Bp.T)).cuda(), n, l, l)
    C = A @ B
    got = oracle.band_to_dense(np.asfortranarray(C.data.cpu().numpy().T), n, 2 * l, 2 * l)
    assert not np.isnan(got).any()
    assert np.max(np.abs(got - oracle.band_to_dense(want, n, 2 * l, 2 * l))) <= tol_rows(np.float64, 2 * l + 1, 1.0, 1.0)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_gbmm_register_kernel(L, oracle, dt):
    """Narrow bands at large n take the register kernel (gbmm_reg.cu); gbmm_force=3 forces it on small, rectangular,
    asymmetric and template-padded shapes (stored rows 1..17 map onto the {3,5,9,17} instantiations)."""
    h = L.handle_for(0)
    cases = [(6, 6, 6, 1, 1, 1, 1), (10, 8, 9, 2, 0, 1, 3), (5, 5, 5, 4, 4, 4, 4), (12, 12, 12, 0, 0, 2, 1), (1, 1, 1, 0, 0, 0, 0),
             (200, 170, 230, 3, 2, 1, 4), (700, 700, 700, 8, 8, 8, 8), (513, 515, 511, 8, 8, 1, 1), (600, 600, 600, 0, 16, 16, 0),
             (1000, 1030, 980, 10, 6, 5, 9), (300, 2000, 300, 2, 2, 2, 2), (2000, 300, 2000, 3, 1, 0, 7), (257, 256, 258, 4, 4, 2, 2),
             (4000, 4000, 4000, 1, 1, 1, 1), (777, 777, 777, 0, 0, 0, 0),
             # exact instantiations for equal odd row counts 7, 11, 13, 15 (bulk-TMA path on the interior tiles)
             (3000, 3000, 3000, 3, 3, 3, 3), (2900, 3100, 3000, 5, 5, 5, 5), (3000, 3000, 3000, 6, 6, 6, 6), (3000, 3000, 3000, 7, 7, 7, 7),
             (3000, 3000, 3000, 2, 4, 5, 1), (3000, 3000, 3000, 10, 0, 3, 7),
             # exact instantiations for equal EVEN row counts 2, 4, 6, 10, 12, 14 (A*A', triangular bands, bidiagonal-like)
             (3000, 3000, 3000, 0, 1, 1, 0), (3000, 3100, 2900, 1, 0, 0, 1), (3000, 3000, 3000, 1, 2, 2, 1), (3000, 3000, 3000, 3, 2, 1, 4),
             (3000, 3000, 3000, 5, 4, 4, 5), (2900, 3000, 3100, 6, 5, 5, 6), (3000, 3000, 3000, 0, 13, 13, 0), (3000, 3000, 3000, 4, 3, 3, 4),
             # no exact instantiation (8 / 16 rows, unequal counts): the run-time-count variant, operands in HBM layout, bulk TMA
             (3000, 3000, 3000, 3, 4, 4, 3), (3100, 2900, 3000, 8, 7, 7, 8), (3000, 3000, 3000, 3, 3, 1, 1), (3000, 3000, 3000, 1, 2, 2, 3),
             (5000, 5000, 5000, 2, 3, 6, 3), (3000, 3000, 3000, 0, 0, 8, 8), (3000, 3000, 3000, 8, 8, 0, 1), (2600, 2700, 2800, 7, 8, 0, 0)]
    for s, (m, n, k, kla, kua, klb, kub) in enumerate(cases):
        A_np, A = _band(L, oracle, m, k, kla, kua, 61 + s, dt)
        B_np, B = _band(L, oracle, k, n, klb, kub, 62 + s, dt)
        want = oracle.gbmm(m, n, k, A_np, kla, kua, B_np, klb, kub)
        h.set_tuning(gbmm_force=3)
        try:
            C = A @ B
            torch.cuda.synchronize()
        finally:
            h.set_tuning(gbmm_force=0)
        tol = tol_rows(NP[dt], min(kla + kua + 1, klb + kub + 1), 1.0, 1.0)
        _cmp_band(oracle, C, want, m, kla + klb, kua + kub, tol)
    # the auto path at a size that fills the machine agrees with the diagonal-wise kernel bit for bit or within tol
    n = 1 << 17
    A_np, A = _band(L, oracle, n, n, 4, 4, 5, dt)
    B_np, B = _band(L, oracle, n, n, 4, 4, 6, dt)
    want = oracle.gbmm(n, n, n, A_np, 4, 4, B_np, 4, 4)
    before = h.launch_count()
    C = A @ B
    assert h.launch_count() - before == 1
    got = np.asfortranarray(C.data.cpu().numpy().T)          # band arrays compared directly (n is too large for dense)
    r = np.arange(17)[:, None]
    i = np.arange(n)[None, :] + r - 8
    inside = (i >= 0) & (i < n)
    assert np.max(np.abs(np.where(inside, got - want, 0.0))) <= tol_rows(NP[dt], 9, 1.0, 1.0)
    assert np.all(got[~inside] == 0.0)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_gbmm_register_kernel_runtime_counts_nan_poisoned(L, oracle, dt):
    """Shapes without an exact instantiation, tight leading dimensions, NaN in every out-of-matrix slot: the run-time-count
    variant (bulk TMA on interior tiles, index-masked cp.async on the edge tiles) and round 1's zero-padded staging
    (gbmm_force = 4) both match the oracle."""
    h = L.handle_for(0)
    for s_, (n, kla, kua, klb, kub) in enumerate([(6000, 4, 3, 3, 4), (6000, 2, 1, 5, 5), (4100, 8, 7, 8, 7), (5000, 0, 5, 9, 0)]):
        A_np = oracle.brand(n, n, kla, kua, seed=90 + s_, dtype=NP[dt])
        B_np = oracle.brand(n, n, klb, kub, seed=95 + s_, dtype=NP[dt])
        want = oracle.gbmm(n, n, n, A_np, kla, kua, B_np, klb, kub)

        def poison(M, l, u):
            P = M.copy()
            r = np.arange(l + u + 1)[:, None]
            i = np.arange(n)[None, :] + r - u
            P[(i < 0) | (i >= n)] = np.nan
            return L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(P.T)).cuda(), n, l, u)

        A, B = poison(A_np, kla, kua), poison(B_np, klb, kub)
        for force in (3, 4):
            h.set_tuning(gbmm_force=force)
            try:
                C = A @ B
                torch.cuda.synchronize()
            finally:
                h.set_tuning(gbmm_force=0)
            got = np.asfortranarray(C.data.cpu().numpy().T)
            l, u = kla + klb, kua + kub
            r = np.arange(l + u + 1)[:, None]
            i = np.arange(n)[None, :] + r - u
            inside = (i >= 0) & (i < n)
            assert not np.isnan(got).any(), (force, kla, kua, klb, kub)
            tol = tol_rows(NP[dt], min(kla + kua + 1, klb + kub + 1), 1.0, 1.0)
            assert np.max(np.abs(np.where(inside, got - want, 0.0))) <= tol, (force, kla, kua, klb, kub)
            assert np.all(got[~inside] == 0.0)


def test_gbmm_register_kernel_alpha_beta_padded_nan(L, oracle):
    """alpha/beta, an output band wider than the product band, padded leading dimensions, and NaNs parked in every
    out-of-matrix slot of A and B (masked by index at staging time)."""
    h = L.handle_for(0)
    m = n = k = 900
    A_np, _ = _band(L, oracle, m, k, 3, 5, 1, torch.float64)
    B_np, _ = _band(L, oracle, k, n, 6, 2, 2, torch.float64)
    C0_np, C0 = _band(L, oracle, m, n, 11, 9, 3, torch.float64)      # band wider than (9,7)
    want = oracle.gbmm(m, n, k, A_np, 3, 5, B_np, 6, 2, klc=11, kuc=9, alpha=-2.0, beta=0.5, Cin=C0_np)

    def poison(M, l, u, rows):
        P = M.copy()
        r = np.arange(l + u + 1)[:, None]
        i = np.arange(P.shape[1])[None, :] + r - u
        P[(i < 0) | (i >= rows)] = np.nan
        return P

    def padded(M, extra):   # (n, lda) tensor with lda > l+u+1, junk in the padding rows
        t = torch.full((M.shape[1], M.shape[0] + extra), float("nan"), dtype=torch.float64, device="cuda")
        t[:, : M.shape[0]] = torch.from_numpy(np.ascontiguousarray(M.T)).cuda()
        return t

    A = L.BandedMatrix(padded(poison(A_np, 3, 5, m), 3), m, 3, 5)
    B = L.BandedMatrix(padded(poison(B_np, 6, 2, k), 1), k, 6, 2)
    h.set_tuning(gbmm_force=3)
    try:
        L.gbmm(A, B, alpha=-2.0, out=C0, beta=0.5)
        torch.cuda.synchronize()
    finally:
        h.set_tuning(gbmm_force=0)
    got = np.asfortranarray(C0.data.cpu().numpy().T)
    dg, dw = oracle.band_to_dense(got, m, 11, 9), oracle.band_to_dense(want, m, 11, 9)
    assert not np.isnan(dg).any()
    assert np.max(np.abs(dg - dw)) <= tol_rows(np.float64, 9, 1.0, 1.0, 2.0, 0.5, 1.0)


def test_gbmm_alpha_beta_and_wider_output_band(L, oracle):
    m = n = k = 500
    A_np, A = _band(L, oracle, m, k, 2, 1, 1, torch.float64)
    B_np, B = _band(L, oracle, k, n, 1, 3, 2, torch.float64)
    C0_np, C0 = _band(L, oracle, m, n, 5, 6, 3, torch.float64)      # band wider than (3,4)
    want = oracle.gbmm(m, n, k, A_np, 2, 1, B_np, 1, 3, klc=5, kuc=6, alpha=2.0, beta=0.5, Cin=C0_np)
    L.gbmm(A, B, alpha=2.0, out=C0, beta=0.5)
    got = np.asfortranarray(C0.data.cpu().numpy().T)
    dg, dw = oracle.band_to_dense(got, m, 5, 6), oracle.band_to_dense(want, m, 5, 6)
    assert np.max(np.abs(dg - dw)) <= tol_rows(np.float64, 4, 1.0, 1.0, 2.0, 0.5, 1.0)
    with pytest.raises(L.LbmArgumentError):                            # output band too narrow
        L.gbmm(A, B, out=L.brand(m, n, 1, 1))
    with pytest.raises(L.DimensionMismatch):
        A @ L.brand(k + 1, n, 1, 1)


def test_applymatrix_config1(L, oracle):
    """BASELINE config 1: ApplyMatrix(*, A, A), A = brand(10_000,10_000,1,1): bandwidths,
    materialize and mul! (README.md:8-26)."""
    n = 10_000
    A_np, A = _band(L, oracle, n, n, 1, 1, 0x1, torch.float64)
    M = L.ApplyMatrix("*", A, A)
    assert M.bandwidths() == (2, 2)
    C = M.materialize()
    want = oracle.gbmm(n, n, n, A_np, 1, 1, A_np, 1, 1)
    _cmp_band(oracle, C, want, n, 2, 2, tol_rows(np.float64, 3, 1.0, 1.0))
    x_np = oracle.fill_uniform(0x9, n)
    y = M @ torch.from_numpy(x_np).cuda()
    t = oracle.gbmv("N", n, n, 1, 1, 1.0, A_np, x_np)
    want_y = oracle.gbmv("N", n, n, 1, 1, 1.0, A_np, t)
    # two stages: stage bounds add (SURVEY.md 8d tolerance row); ||A x|| <= 3
    assert np.max(np.abs(y.cpu().numpy() - want_y)) <= 2 * tol_rows(np.float64, 3, 1.0, 3.0)
    # materialised product applied to x agrees with the lazy apply
    y2 = C @ torch.from_numpy(x_np).cuda()
    assert np.max(np.abs(y2.cpu().numpy() - want_y)) <= 2 * tol_rows(np.float64, 5, 3.0, 1.0)
    # three-factor chain, mixed kinds
    T = L.SymTridiagonal(torch.from_numpy(oracle.fill_uniform(3, n)).cuda(), torch.from_numpy(oracle.fill_uniform(4, n - 1)).cuda())
    M3 = L.ApplyMatrix("*", A, T, A.T)
    y3 = M3 @ torch.from_numpy(x_np).cuda()
    dense_y = A.to_dense() @ (T.to_dense() @ (A.to_dense().t() @ torch.from_numpy(x_np).cuda()))
    assert torch.max(torch.abs(y3 - dense_y)).item() <= 1e-12


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_applymatrix_chain_apply_one_call(L, oracle, dt):
    """mul!(y, ApplyMatrix(*, F0, F1, F2), x, alpha, beta) through lbm_?gbmv_chain: rectangular factors, different bands,
    alpha/beta, temporaries in the handle's scratch (one C call, nf launches)."""
    h = L.handle_for(0)
    shapes = [(700, 900, 2, 1), (900, 650, 0, 3), (650, 800, 4, 4)]
    nps, devs = [], []
    for s_, (m, n, l, u) in enumerate(shapes):
        a_np, a = _band(L, oracle, m, n, l, u, 21 + s_, dt)
        nps.append(a_np)
        devs.append(a)
    x = oracle.fill_uniform(9, 800, NP[dt])
    y0 = oracle.fill_uniform(10, 700, NP[dt])
    v = x
    for (m, n, l, u), a_np in zip(reversed(shapes[1:]), reversed(nps[1:])):
        v = oracle.gbmv("N", m, n, l, u, 1.0, a_np, v)
    m, n, l, u = shapes[0]
    want = oracle.gbmv("N", m, n, l, u, -1.5, nps[0], v, 0.5, y0.copy())
    M = L.ApplyMatrix("*", *devs)
    y = torch.from_numpy(y0.copy()).cuda()
    before = h.launch_count()
    M.mul_(y, torch.from_numpy(x).cuda(), -1.5, 0.5)
    assert h.launch_count() - before == 3
    # stage bounds add: ||F2 x|| <= 9, ||F1 F2 x|| <= 36
    tol = tol_rows(NP[dt], 9, 1.0, 1.0) + tol_rows(NP[dt], 4, 1.0, 9.0) + tol_rows(NP[dt], 4, 1.0, 36.0, 1.5, 0.5, 1.0)
    assert np.max(np.abs(y.cpu().numpy() - want)) <= tol
    with pytest.raises(L.LbmArgumentError):          # inner sizes must chain (C ABI: argument 4)
        bad = L.ApplyMatrix.__new__(L.ApplyMatrix)
        bad.f, bad.args = "*", (devs[0], devs[2])
        bad.mul_(torch.empty(700, dtype=dt, device="cuda"), torch.from_numpy(x).cuda())


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_chain3_batched(L, oracle, dt):
    npdt = NP[dt]
    for (batch, n, kl, ku) in [(5, 300, [4, 4, 4], [4, 4, 4]), (3, 1000, [1, 2, 0], [0, 1, 3]), (2, 70, [0, 0, 0], [0, 0, 0]),
                               (7, 4096, [4, 4, 4], [4, 4, 4]), (3, 1000, [1, 1, 1], [1, 1, 1]), (2, 513, [2, 2, 2], [2, 2, 2]),
                               (4, 37, [4, 4, 4], [4, 4, 4]), (2, 1027, [4, 4, 4], [4, 4, 4]), (1, 3, [4, 4, 4], [4, 4, 4]),
                               (3, 1024, [2, 2, 2], [2, 2, 2]), (2, 516, [1, 1, 1], [1, 1, 1]), (2, 8, [4, 4, 4], [4, 4, 4]), (2, 4, [1, 1, 1], [1, 1, 1])]:
        fs = [oracle.fill_uniform(1 + i, batch * n * (kl[i] + ku[i] + 1), npdt).reshape(batch, n, kl[i] + ku[i] + 1)
              for i in range(3)]
        want = oracle.gbmm_chain3_batched(n, kl, ku, *fs)
        got = L.chain3_batched(*[torch.from_numpy(f).cuda() for f in fs], kl, ku).cpu().numpy()
        for variant in ((1, 3) if dt == torch.float32 else ()):   # split-column / transposed-smem variants: same in-matrix entries
            hh = L.handle_for(0)
            hh.set_tuning(chain_variant=variant)
            try:
                got3 = L.chain3_batched(*[torch.from_numpy(f).cuda() for f in fs], kl, ku).cpu().numpy()
            finally:
                hh.set_tuning(chain_variant=0)
            for b in range(batch):
                assert np.max(np.abs(oracle.band_to_dense(np.asfortranarray(got3[b].T), n, sum(kl), sum(ku))
                                     - oracle.band_to_dense(np.asfortranarray(want[b].T), n, sum(kl), sum(ku)))) <= \
                    tol_rows(npdt, kl[0] + ku[0] + 1, 1.0, 1.0) * (kl[2] + ku[2] + 1) + tol_rows(npdt, kl[2] + ku[2] + 1, float(kl[0] + ku[0] + 1), 1.0)
        nb = [kl[i] + ku[i] + 1 for i in range(3)]
        # per-stage bounds: stage 1 terms <= nb0 (entries <= 1), stage 2 terms <= nb2 with |AB| <= nb0
        tol = tol_rows(npdt, nb[0], 1.0, 1.0) * nb[2] + tol_rows(npdt, nb[2], float(nb[0]), 1.0)
        assert got.shape == want.shape
        ld = want.shape[2]
        l3, u3 = sum(kl), sum(ku)
        for b in range(batch):   # compare in-matrix entries
            dg = oracle.band_to_dense(np.asfortranarray(got[b].T), n, l3, u3)
            dw = oracle.band_to_dense(np.asfortranarray(want[b].T), n, l3, u3)
            assert np.max(np.abs(dg - dw)) <= tol, (batch, n, kl, ku)
        assert ld == l3 + u3 + 1


def test_chain3_register_kernel_vs_generic_and_nan_corners(L, oracle):
    """The register-blocked fused chain == the generic fused chain == oracle, and NaNs parked in
    out-of-matrix storage slots never reach an in-matrix entry."""
    h = L.handle_for(0)
    batch, n, kl, ku = 3, 1500, [4, 4, 4], [4, 4, 4]
    fs = [oracle.fill_uniform(1 + i, batch * n * 9, np.float32).reshape(batch, n, 9).copy() for i in range(3)]
    want = oracle.gbmm_chain3_batched(n, kl, ku, *fs)
    for f in fs:                                       # poison every out-of-matrix slot
        for j in list(range(4)) + list(range(n - 4, n)):
            for r in range(9):
                i = j + r - 4
                if i < 0 or i >= n:
                    f[:, j, r] = np.nan
    gs = [torch.from_numpy(f).cuda() for f in fs]
    tol = tol_rows(np.float32, 9, 1.0, 1.0) * 9 + tol_rows(np.float32, 9, 9.0, 1.0)
    outs = []
    for force, variant in ((0, 0), (0, 1), (0, 2), (0, 3), (1, 0)):     # warp-autonomous (default), split-column, transposed x2, generic
        h.set_tuning(chain_force_generic=force, chain_variant=variant)
        try:
            outs.append(L.chain3_batched(*gs, kl, ku).cpu().numpy())
        finally:
            h.set_tuning(chain_force_generic=0, chain_variant=0)
    for got in outs:
        for b in range(batch):
            dg = oracle.band_to_dense(np.asfortranarray(got[b].T), n, 12, 12)
            dw = oracle.band_to_dense(np.asfortranarray(want[b].T), n, 12, 12)
            assert not np.isnan(dg).any()
            assert np.max(np.abs(dg - dw)) <= tol


def test_chain3_warp_kernel_persistent_loop(L, oracle):
    """More tiles than resident warps (148 SMs x 8 warps): every warp walks several tiles, so the
    mbarrier parities, the A*B/D slab aliasing and the next-tile prefetch are all exercised; ragged
    last tiles (n/4 not a multiple of the tile's 30 column groups); out-of-matrix slots of D are 0."""
    for (batch, n, k) in ((96, 4096, 4), (300, 1000, 4), (700, 516, 2), (1500, 128, 1), (40, 4, 1)):
        kl = ku = [k, k, k]
        ld = 2 * k + 1
        fs = [oracle.fill_uniform(11 + i, batch * n * ld, np.float32).reshape(batch, n, ld) for i in range(3)]
        gs = [torch.from_numpy(f).cuda() for f in fs]
        got = L.chain3_batched(*gs, kl, ku).cpu().numpy()          # default: single-buffered A/B
        h = L.handle_for(0)
        h.set_tuning(chain_variant=5)                               # A/B double-buffered, a tile ahead
        try:
            got5 = L.chain3_batched(*gs, kl, ku).cpu().numpy()
        finally:
            h.set_tuning(chain_variant=0)
        assert np.array_equal(got, got5)                            # same arithmetic, only the copy schedule differs
        h.set_tuning(chain_variant=1)
        try:
            ref_gpu = L.chain3_batched(*gs, kl, ku).cpu().numpy()
        finally:
            h.set_tuning(chain_variant=0)
        tol = tol_rows(np.float32, ld, 1.0, 1.0) * ld + tol_rows(np.float32, ld, float(ld), 1.0)
        # the split-column kernel also stores 0 in out-of-matrix slots: whole arrays are comparable
        assert np.max(np.abs(got - ref_gpu)) <= tol, (batch, n, k)
        for b in (0, batch // 2, batch - 1):
            want = oracle.gbmm_chain3_batched(n, kl, ku, *[f[b:b + 1] for f in fs])[0]
            dg = oracle.band_to_dense(np.asfortranarray(got[b].T), n, 3 * k, 3 * k)
            dw = oracle.band_to_dense(np.asfortranarray(want.T), n, 3 * k, 3 * k)
            assert np.max(np.abs(dg - dw)) <= tol, (batch, n, k, b)
        j = np.arange(n)[:, None]
        i = j + np.arange(6 * k + 1)[None, :] - 3 * k
        assert np.all(got[:, (i < 0) | (i >= n)] == 0.0)


def _lap(L, n, dtype=torch.float64):
    h = 1.0 / n
    return L.banded_from_diagonals(n, {0: -2.0 / h**2, 1: 1.0 / h**2, -1: 1.0 / h**2}, dtype=dtype)


def test_kronsum_laplacian(L, oracle):
    """2-D Laplacian BlockKron(D2,I)+BlockKron(I,D2) (test/test_blockkron.jl:296-306)."""
    for n in (4, 33, 257, 1000, 1024, 2048):
        D = _lap(L, n)
        Dxx = L.BlockKron(D, L.Eye(n))
        Dyy = L.BlockKron(L.Eye(n), D)
        Lap = L.KronSum.from_broadcast(Dxx, Dyy)
        assert Lap.blockbandwidths() == (1, 1) and Lap.subblockbandwidths() == (1, 1)
        x_np = oracle.fill_uniform(5, n * n)
        D_np = np.asfortranarray(D.data.cpu().numpy().T)
        want = oracle.kronsum2d_mv(n, n, D_np, 1, 1, D_np, 1, 1, x_np)
        got = (Lap @ torch.from_numpy(x_np).cuda()).cpu().numpy()
        scale = 2.0 * n * n                      # ||D2||_max = 2/h^2
        assert np.max(np.abs(got - want)) <= tol_rows(np.float64, 6, scale, 1.0)
        if n <= 33:                              # and against the literally formed Kronecker matrices
            Dd = oracle.band_to_dense(D_np, n, 1, 1)
            dense = (np.kron(Dd, np.eye(n)) + np.kron(np.eye(n), Dd)) @ x_np
            assert np.max(np.abs(got - dense)) <= tol_rows(np.float64, 6, scale, 1.0)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_kronsum_general_bands_rectangular_grid(L, oracle, dt):
    npdt = NP[dt]
    for (n1, n2, la, ua, lb, ub) in [(7, 5, 1, 1, 1, 1), (300, 200, 2, 1, 0, 3), (64, 1000, 3, 3, 2, 2), (1, 50, 0, 0, 1, 1),
                                     (50, 1, 1, 1, 0, 0), (513, 511, 1, 1, 1, 1),
                                     # TMA-ring kernel shapes (n2 % 4 == 0, n2 >= 512): tri x tri, diag x tri,
                                     # tri x diag, penta x penta, mixed, strips/chunks that do not divide the grid
                                     (100, 1024, 1, 1, 1, 1), (37, 516, 0, 0, 1, 1), (300, 512, 1, 1, 0, 0),
                                     (129, 2052, 2, 2, 2, 2), (77, 1000, 1, 1, 2, 2), (9, 768, 2, 2, 1, 1),
                                     (1030, 520, 0, 2, 2, 0), (16, 4096, 1, 1, 1, 1)]:
        A_np, A = _band(L, oracle, n1, n1, la, ua, 3, dt)
        B_np, B = _band(L, oracle, n2, n2, lb, ub, 4, dt)
        x_np = oracle.fill_uniform(5, n1 * n2, npdt)
        y0 = oracle.fill_uniform(6, n1 * n2, npdt)
        want = oracle.kronsum2d_mv(n1, n2, A_np, la, ua, B_np, lb, ub, x_np, 1.5, 0.5, y0)
        y = torch.from_numpy(y0.copy()).cuda()
        L.KronSum(A, B).mul_(y, torch.from_numpy(x_np).cuda(), 1.5, 0.5)
        tol = tol_rows(npdt, la + ua + lb + ub + 2, 1.0, 1.0, 1.5, 0.5, 1.0)
        assert np.max(np.abs(y.cpu().numpy() - want)) <= tol


def test_kronsum_ring_tunings_match_tile_kernel(L, oracle):
    """TMA-ring Kronecker kernel under several launch shapes == generic tile kernel == oracle."""
    h = L.handle_for(0)
    n1, n2 = 203, 2048
    A_np, A = _band(L, oracle, n1, n1, 1, 1, 3, torch.float64)
    B_np, B = _band(L, oracle, n2, n2, 1, 1, 4, torch.float64)
    x_np = oracle.fill_uniform(5, n1 * n2)
    want = oracle.kronsum2d_mv(n1, n2, A_np, 1, 1, B_np, 1, 1, x_np)
    x = torch.from_numpy(x_np).cuda()
    tol = tol_rows(np.float64, 6, 1.0, 1.0)
    for tune in ({"kron_force_tile": 1}, {"kron_tr": 256, "kron_tc": 4, "kron_stages": 1, "kron_ctas_per_sm": 1},
                 {"kron_tr": 512, "kron_tc": 16, "kron_stages": 3}, {"kron_tr": 1024, "kron_tc": 8, "kron_stages": 2},
                 {"kron_tr": 256, "kron_tc": 64, "kron_stages": 2, "kron_ctas_per_sm": 2}):
        h.set_tuning(**tune)
        try:
            y = torch.full((n1 * n2,), float("nan"), dtype=torch.float64, device="cuda")
            L.KronSum(A, B).mul_(y, x)
            got = y.cpu().numpy()
        finally:
            h.set_tuning(**{k: 0 for k in tune})
        assert np.max(np.abs(got - want)) <= tol, tune


def test_blockkron_apply(L, oracle):
    """BlockKron(A,B)*x = vec(B X A') without forming kron(A,B) (src/blockkron.jl:440 identity)."""
    # the last three shapes are large enough for the streaming ring kernel (each pass = Kronecker sum with a zero partner)
    for (n1, n2, la, ua, lb, ub) in [(4, 4, 1, 1, 0, 0), (40, 30, 1, 1, 2, 1), (257, 300, 2, 2, 1, 1), (300, 1024, 1, 1, 1, 1),
                                     (129, 2050, 2, 2, 1, 1), (64, 512, 1, 1, 3, 0)]:
        A_np, A = _band(L, oracle, n1, n1, la, ua, 3, torch.float64)
        B_np, B = _band(L, oracle, n2, n2, lb, ub, 4, torch.float64)
        K = L.BlockKron(A, B)
        assert K.blockbandwidths() == (la, ua) and K.subblockbandwidths() == (lb, ub)
        x_np = oracle.fill_uniform(5, n1 * n2)
        want = oracle.kron2d_mv(n1, n2, A_np, la, ua, B_np, lb, ub, x_np)
        got = (K @ torch.from_numpy(x_np).cuda()).cpu().numpy()
        tol = tol_rows(np.float64, (la + ua + 1) * (lb + ub + 1), 1.0, 1.0)
        assert np.max(np.abs(got - want)) <= tol
        if n1 * n2 <= 2000:
            dense = np.kron(oracle.band_to_dense(A_np, n1, la, ua), oracle.band_to_dense(B_np, n2, lb, ub)) @ x_np
            assert np.max(np.abs(got - dense)) <= tol


def _dense_banded(L, M):
    """Dense numpy matrix -> full-band BandedMatrix on the device."""
    m, n = M.shape
    data = np.zeros((m + n - 1, n), order="F")
    l, u = m - 1, n - 1
    for j in range(n):
        for i in range(m):
            data[u + i - j, j] = M[i, j]
    return L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(data.T)).cuda(), m, l, u)


def test_krontrav_apply_reference_kats(L, oracle, kats):
    """The reference's own KATs for KronTrav*DiagTrav, computed on the GPU without forming K."""
    k = next(k for k in kats if k["kind"] == "krontrav_mul")            # test/test_blockkron.jl:139-148
    A, B, X, Y = (np.array(k[s], dtype=np.float64) for s in "ABXY")
    K = L.KronTrav(_dense_banded(L, A), _dense_banded(L, B))
    for Z in (X, Y):                                                       # K*DiagTrav(X) == K*DiagTrav(Y)
        got = K.mul_diagtrav(L.DiagTrav(torch.from_numpy(Z).cuda())).vector().cpu().tolist()
        assert got == [211.0, 291.0, 533.0]
    for k in [k for k in kats if k["kind"] == "diagtrav"]:                 # DiagTrav index map on the device
        Xn = np.array(k["X"], dtype=np.float64)
        if max(Xn.shape) == min(Xn.shape) or True:
            got = L.DiagTrav(torch.from_numpy(Xn).cuda()).vector().cpu().tolist()
            want = oracle.diagtrav(oracle.zero_bottomright(Xn.copy())).tolist()
            assert got == want, k["src"]
    k = next(k for k in kats if k["kind"] == "krontrav3_matrix")           # explicit 4x4 KronTrav(A,B,C), :154-164
    A, B, Cm = (np.array(k[s], dtype=np.float64) for s in "ABC")
    K3 = L.KronTrav(_dense_banded(L, A), _dense_banded(L, B), _dense_banded(L, Cm))
    n = 2
    order = oracle.diagtrav(np.arange(n**3).reshape(n, n, n))
    cols = []
    for flat in order:
        Xt = np.zeros((n, n, n))
        Xt[np.unravel_index(int(flat), (n, n, n))] = 1.0
        cols.append(K3.mul_diagtrav(L.DiagTrav(torch.from_numpy(Xt).cuda())).vector().cpu().numpy())
    assert np.stack(cols, axis=1).astype(int).tolist() == k["expect"]
    for k in [k for k in kats if k["kind"] == "krontrav_laplace2d"]:       # A*DiagTrav(X) == DiagTrav(X*D'), :201-207
        n = k["n"]
        D = L.banded_from_diagonals(n, {0: float(k["diag"]), 1: float(k["off"]), -1: float(k["off"])})
        Dd = D.to_dense().cpu().numpy()
        Xn = np.triu(np.random.default_rng(0).standard_normal((n, n)))[:, ::-1].copy()
        Xg = L.DiagTrav(torch.from_numpy(Xn).cuda())
        if "X * Δ'" in k["snippet"]:
            got = L.KronTrav(D, L.Eye(n)).mul_diagtrav(Xg).vector().cpu().numpy()
            want = oracle.diagtrav(Xn @ Dd.T)
        else:
            got = L.KronTrav(L.Eye(n), D).mul_diagtrav(Xg).vector().cpu().numpy()
            want = oracle.diagtrav(Dd @ Xn)
        assert np.array_equal(got, want), k["src"]                        # exact for the +-1/-2 stencil


def test_kron3d_mode_products(L, oracle):
    """lbm_dkron3d_mv == the three mode-product sweeps of src/blockkron.jl:441-450 (numpy einsum),
    and KronTrav(A,B,C)*DiagTrav(X) == the oracle's restatement including the DiagTrav ordering."""
    rng = np.random.default_rng(2)
    for n, bw in ((4, (1, 1)), (9, (2, 1)), (33, (1, 1)), (64, (3, 3))):
        facs = [_band(L, oracle, n, n, bw[0], bw[1], 40 + i, torch.float64) for i in range(3)]
        dA, dB, dC = (oracle.band_to_dense(f[0], n, *bw) for f in facs)
        X = rng.standard_normal((n, n, n))
        K3 = L.KronTrav(facs[0][1], facs[1][1], facs[2][1])
        out = K3.mul_diagtrav(L.DiagTrav(torch.from_numpy(X).cuda()))
        want_vec = oracle.krontrav3_mul_diagtrav(dA, dB, dC, X)
        tol = tol_rows(np.float64, (bw[0] + bw[1] + 1) ** 3, 1.0, float(np.abs(X).max()))
        assert np.max(np.abs(out.vector().cpu().numpy() - want_vec)) <= tol


def test_broadcast_sum_of_banded(L, oracle):
    """BroadcastArray(+/-, banded...)*x == A*x +/- B*x (union band), incl. mixed structured kinds and
    a lazy product inside the sum (test/test_blockkron.jl:345-346 pattern)."""
    n = 5000
    A_np, A = _band(L, oracle, n, n, 2, 1, 1, torch.float64)
    B_np, B = _band(L, oracle, n, n, 0, 3, 2, torch.float64)
    d, e = oracle.fill_uniform(3, n), oracle.fill_uniform(4, n - 1)
    S = L.SymTridiagonal(torch.from_numpy(d).cuda(), torch.from_numpy(e).cuda())
    x_np = oracle.fill_uniform(9, n)
    x = torch.from_numpy(x_np).cuda()
    Ax = oracle.gbmv("N", n, n, 2, 1, 1.0, A_np, x_np)
    Bx = oracle.gbmv("N", n, n, 0, 3, 1.0, B_np, x_np)
    Sx = oracle.tridiag_mv(n, e, d, e, x_np)
    M = L.BroadcastMatrix("+", A, B, S)
    assert M.bandwidths() == (2, 3)
    assert np.max(np.abs((M @ x).cpu().numpy() - (Ax + Bx + Sx))) <= tol_rows(np.float64, 4 + 4 + 3, 1.0, 1.0)
    Mm = L.BroadcastMatrix("-", A, B)
    y0 = oracle.fill_uniform(5, n)
    y = torch.from_numpy(y0.copy()).cuda()
    Mm.mul_(y, x, 2.0, 0.5)
    assert np.max(np.abs(y.cpu().numpy() - (2.0 * (Ax - Bx) + 0.5 * y0))) <= tol_rows(np.float64, 8, 1.0, 1.0, 2.0, 0.5, 1.0)
    P = L.BroadcastMatrix("+", L.ApplyMatrix("*", A, B), S)          # (A*B + S) x
    want = oracle.gbmv("N", n, n, 2, 1, 1.0, A_np, Bx) + Sx
    assert np.max(np.abs((P @ x).cpu().numpy() - want)) <= 2 * tol_rows(np.float64, 4, 1.0, 4.0) + tol_rows(np.float64, 3, 1.0, 1.0)
    with pytest.raises(L.DimensionMismatch):
        L.BroadcastMatrix("+", A, L.brand(n + 1, n + 1, 1, 1))


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_broadcast_sum_fused_gbmv2(L, oracle, dt):
    """mul!(y, BroadcastArray(+/-, A1, A2), x) through the fused lbm_?gbmv2 (one pass over x and y) equals
    alpha*(A1 x +/- A2 x) + beta*y from two oracle gbmv's; rectangular, different bands, wide-band fallback,
    NaNs in out-of-matrix slots, and exactly ONE kernel launch on the fused path."""
    npdt = NP[dt]
    h = L.handle_for(0)
    cases = [  # (m, n, (kl1, ku1), (kl2, ku2), fused?)
        (5000, 5000, (2, 1), (0, 3), True), (100003, 100003, (1, 1), (1, 1), True), (3000, 3100, (8, 8), (1, 0), True),
        (3100, 3000, (0, 0), (4, 2), True), (7, 7, (1, 1), (2, 2), True), (4000, 4000, (16, 16), (16, 16), False),   # f64 stage > 100 KB: two launches
        (2000, 2000, (128, 128), (1, 1), False), (1, 1, (0, 0), (0, 0), True),
    ]
    for s, (m, n, (kl1, ku1), (kl2, ku2), fused) in enumerate(cases):
        mats = []
        for f, (kl, ku) in enumerate(((kl1, ku1), (kl2, ku2))):
            lda = kl + ku + 1
            A_np = np.asfortranarray(oracle.fill_uniform(500 + 2 * s + f, lda * n, npdt).reshape((lda, n), order="F"))
            Ap = A_np.copy()
            i = np.arange(n)[None, :] + np.arange(lda)[:, None] - ku
            Ap[(i < 0) | (i >= m)] = np.nan
            mats.append((A_np, L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(Ap.T)).cuda(), m, kl, ku)))
        x_np = oracle.fill_uniform(600 + s, n, npdt)
        y_np = oracle.fill_uniform(700 + s, m, npdt)
        x = torch.from_numpy(x_np).cuda()
        for op, sg in (("+", 1.0), ("-", -1.0)):
            for alpha, beta in ((1.0, 0.0), (2.0, 0.5)):
                y = torch.from_numpy(y_np.copy()).cuda()
                before = h.launch_count()
                L.BroadcastMatrix(op, mats[0][1], mats[1][1]).mul_(y, x, alpha, beta)
                torch.cuda.synchronize()
                if fused:
                    assert h.launch_count() - before == 1
                want = (oracle.gbmv("N", m, n, kl1, ku1, alpha, mats[0][0], x_np, beta, y_np)
                        + oracle.gbmv("N", m, n, kl2, ku2, sg * alpha, mats[1][0], x_np))
                tol = tol_rows(npdt, kl1 + ku1 + 1 + kl2 + ku2 + 1, 1.0, 1.0, alpha, beta, 1.0)
                assert np.max(np.abs(y.cpu().numpy() - want)) <= tol, (m, n, kl1, ku1, kl2, ku2, op, alpha, beta)
