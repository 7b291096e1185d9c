"""GPU parity: banded mul! (lbm_?gbmv through the C ABI) against the oracle on the same
seeded inputs.  Tolerance (north_star): |y_gpu - y_ref| <= 4*eps*(l+u+1)*||A||*||x|| per row."""
import numpy as np
import pytest
import torch

from tests._util import tol_rows

pytestmark = pytest.mark.gpu

NP = {torch.float64: np.float64, torch.float32: np.float32}

SHAPES = [  # (m, n, kl, ku)
    (1, 1, 0, 0), (5, 5, 1, 1), (7, 7, 0, 2), (9, 9, 3, 0), (64, 64, 8, 8), (33, 40, 2, 5), (40, 33, 5, 2),
    (3, 3, 4, 4), (2, 9, 1, 1), (9, 2, 1, 1), (10, 4, 0, 0), (1000, 1000, 1, 1), (4097, 4097, 1, 1),
    (5000, 5003, 8, 8), (3001, 2999, 4, 2), (777, 777, 16, 16), (2000, 2000, 20, 3), (1500, 1500, 40, 40),
    (3000, 3000, 128, 128), (2500, 2600, 100, 30), (600, 600, 300, 300), (100000, 100000, 1, 1),
    (70001, 70001, 8, 8), (20000, 20000, 128, 128), (50, 3000, 2, 2), (3000, 50, 2, 2),
]


def _run(L, oracle, dt, trans, m, n, kl, ku, alpha, beta, seed, lda_pad=0, tune=None):
    npdt = NP[dt]
    lda = kl + ku + 1 + lda_pad
    A_np = np.asfortranarray(oracle.fill_uniform(seed, lda * n, npdt).reshape((lda, n), order="F"))
    nin, nout = (n, m) if trans == "N" else (m, n)
    x_np = oracle.fill_uniform(seed + 1, nin, npdt)
    y_np = oracle.fill_uniform(seed + 2, nout, npdt)
    want = oracle.gbmv(trans, m, n, kl, ku, alpha, A_np, x_np, beta, y_np, lda=lda)
    A = L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(A_np.T)).cuda(), m, kl, ku)
    assert A.lda == lda
    x = torch.from_numpy(x_np).cuda()
    y = torch.from_numpy(y_np.copy()).cuda()
    h = L.handle_for(0)
    if tune:
        h.set_tuning(**tune)
    try:
        L.mul_(y, A if trans == "N" else A.T, x, alpha, beta)
        torch.cuda.synchronize()
    finally:
        if tune:
            h.set_tuning(**{k: 0 for k in tune})
    err = np.max(np.abs(y.cpu().numpy() - want)) if nout else 0.0
    tol = tol_rows(npdt, kl + ku + 1, 1.0, 1.0, alpha, beta, 1.0)
    assert err <= tol, (m, n, kl, ku, trans, alpha, beta, err, tol)
    return err / tol


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("trans", ["N", "T"])
def test_gbmv_shapes(L, oracle, dt, trans):
    worst = 0.0
    for s, (m, n, kl, ku) in enumerate(SHAPES):
        for alpha, beta in [(1.0, 0.0), (2.0, 1.0)]:
            worst = max(worst, _run(L, oracle, dt, trans, m, n, kl, ku, alpha, beta, 100 + 7 * s))
    print(f"max err/tol = {worst:.3f}")


@pytest.mark.parametrize("force", [1, 2, 3, 4])
def test_gbmv_every_kernel_path(L, oracle, force):
    """direct (1), streamed-tile (2), wide column-streaming (3) and the non-ring wide-T warp kernel (4)
    on the same inputs."""
    shapes = [(5000, 5000, 1, 1), (4100, 4100, 8, 8), (9000, 9100, 16, 16), (3000, 3000, 2, 7)]
    if force == 4:
        shapes = [(3000, 3100, 40, 40), (2000, 1900, 128, 128), (700, 700, 300, 20)]
    for (m, n, kl, ku) in shapes:
        for trans in ("N", "T"):
            _run(L, oracle, torch.float64, trans, m, n, kl, ku, 1.5, 0.5, 31, tune={"gbmv_force": force})


@pytest.mark.parametrize("tune", [
    {"gbmv_tile": 256, "gbmv_stages": 1, "gbmv_ctas_per_sm": 1},
    {"gbmv_tile": 512, "gbmv_stages": 4, "gbmv_ctas_per_sm": 2},
    {"gbmv_tile": 2048, "gbmv_stages": 2, "gbmv_ctas_per_sm": 4},
    {"gbmvw_chunk": 32, "gbmvw_stages": 2, "gbmv_force": 3},
    {"gbmvw_chunk": 64, "gbmvw_stages": 4, "gbmv_force": 3},
])
def test_gbmv_tunings(L, oracle, tune):
    for (m, n, kl, ku) in [(300001, 300001, 1, 1), (120000, 120000, 8, 8), (50000, 50000, 33, 31)]:
        _run(L, oracle, torch.float64, "N", m, n, kl, ku, 1.0, 0.0, 5, tune=tune)
        _run(L, oracle, torch.float32, "N", m, n, kl, ku, 1.0, 0.0, 5, tune=tune)


def test_gbmv_padded_lda_and_unaligned(L, oracle):
    _run(L, oracle, torch.float64, "N", 4000, 4000, 2, 2, 1.0, 0.0, 3, lda_pad=3)
    _run(L, oracle, torch.float64, "T", 4000, 4000, 2, 2, 1.0, 0.0, 3, lda_pad=1)
    # unaligned base pointers (views at an odd element offset) take the direct kernel
    n, kl, ku = 3001, 1, 1
    buf = torch.empty(3 * n + 1, dtype=torch.float64, device="cuda")
    L.fill_uniform(buf, 11)
    data = buf[1:].view(n, 3)
    A = L.BandedMatrix(data, n, kl, ku)
    xb = torch.empty(n + 1, dtype=torch.float64, device="cuda")
    L.fill_uniform(xb, 12)
    x = xb[1:]
    y = A @ x
    want = oracle.gbmv("N", n, n, kl, ku, 1.0, np.asfortranarray(data.cpu().numpy().T), x.cpu().numpy())
    assert np.max(np.abs(y.cpu().numpy() - want)) <= tol_rows(np.float64, 3, 1.0, 1.0)


def test_gbmv_wide_padded_lda_rectangular(L, oracle):
    """Wide-band kernels (column-streaming N, ring T, warp T) with lda > kl+ku+1 and m != n."""
    for (m, n, kl, ku, pad) in [(3000, 3100, 70, 60, 2), (2500, 2400, 128, 128, 7), (900, 1500, 200, 10, 1), (1500, 900, 10, 200, 4)]:
        for trans in ("N", "T"):
            for dt in (torch.float64, torch.float32):
                _run(L, oracle, dt, trans, m, n, kl, ku, 1.25, 0.5, 77, lda_pad=pad)
    _run(L, oracle, torch.float64, "T", 3000, 3100, 70, 60, 1.0, 0.0, 78, lda_pad=3, tune={"gbmv_force": 4})
    # tall / wide matrices whose trailing rows (columns) see no band at all: y <- beta*y there
    _run(L, oracle, torch.float64, "N", 5000, 100, 40, 40, 2.0, 0.5, 79)
    _run(L, oracle, torch.float64, "T", 100, 5000, 40, 40, 2.0, 0.5, 79)
    _run(L, oracle, torch.float64, "N", 5000, 100, 2, 2, 2.0, 0.5, 80)
    _run(L, oracle, torch.float64, "T", 100, 5000, 2, 2, 2.0, 0.5, 80)


def test_gbmv_masks_by_index_not_value(L, oracle):
    """Out-of-matrix storage slots hold NaN, y holds NaN and beta == 0: nothing may leak."""
    for (m, n, kl, ku) in [(3000, 3000, 2, 3), (2000, 2000, 128, 128), (700, 900, 5, 1)]:
        A_np = oracle.brand(m, n, kl, ku, seed=5)
        clean = oracle.gbmv("N", m, n, kl, ku, 1.0, A_np, np.ones(n))
        for j in list(range(min(n, ku + 1))) + list(range(max(0, m - kl - 1 - max(0, m - n)), n)):
            for r in range(kl + ku + 1):
                i = j + r - ku
                if i < 0 or i >= m:
                    A_np[r, j] = np.nan
        A = L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(A_np.T)).cuda(), m, kl, ku)
        y = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
        L.mul_(y, A, torch.ones(n, dtype=torch.float64, device="cuda"))
        got = y.cpu().numpy()
        assert not np.isnan(got).any()
        assert np.max(np.abs(got - clean)) <= tol_rows(np.float64, kl + ku + 1, 1.0, 1.0)


def test_gbmv_generator_matches_oracle(L, oracle):
    for dt in (torch.float64, torch.float32):
        t = torch.empty(100003, dtype=dt, device="cuda")
        L.fill_uniform(t, 0x1, offset=17)
        assert np.array_equal(t.cpu().numpy(), oracle.fill_uniform(0x1, 100003, NP[dt], offset=17))


def test_gbmv_dimension_mismatch(L):
    A = L.brand(10, 12, 1, 1)
    with pytest.raises(L.DimensionMismatch):
        L.mul_(torch.zeros(10, dtype=torch.float64, device="cuda"), A, torch.zeros(11, dtype=torch.float64, device="cuda"))
    with pytest.raises(L.DimensionMismatch):
        L.mul_(torch.zeros(9, dtype=torch.float64, device="cuda"), A, torch.zeros(12, dtype=torch.float64, device="cuda"))


def test_gbmv_host_entry_point(L, oracle):
    """lbm_dgbmv_host: the e2e call (host buffers in, host buffer out)."""
    import ctypes as C
    m = n = 50001
    kl = ku = 8
    A_np = oracle.brand(m, n, kl, ku, seed=21)
    x_np = oracle.fill_uniform(22, n)
    y_np = np.zeros(m)
    h = L.handle_for(0)
    rc = L.lib().lbm_dgbmv_host(h.ptr, b"N", m, n, kl, ku, 1.0, A_np.ctypes.data_as(C.c_void_p), kl + ku + 1,
                                x_np.ctypes.data_as(C.c_void_p), 0.0, y_np.ctypes.data_as(C.c_void_p))
    assert rc == 0
    want = oracle.gbmv("N", m, n, kl, ku, 1.0, A_np, x_np)
    assert np.max(np.abs(y_np - want)) <= tol_rows(np.float64, kl + ku + 1, 1.0, 1.0)


def test_gbmv_host_entry_point_pipelined(L, oracle):
    """Large host operands take the 3-stream chunked pipeline (H2D of A overlapped with the
    sub-slab kernels and the D2H of y); result identical to the oracle's full sweep."""
    import ctypes as C
    for (n, kl, ku, beta) in (((1 << 25) + 3, 1, 1, 0.0), (12_000_001, 2, 3, 0.5)):
        A_np = oracle.brand(n, n, kl, ku, seed=21)
        x_np = oracle.fill_uniform(22, n)
        y0 = oracle.fill_uniform(23, n)
        y_np = y0.copy()
        h = L.handle_for(0)
        rc = L.lib().lbm_dgbmv_host(h.ptr, b"N", n, n, kl, ku, 1.5, A_np.ctypes.data_as(C.c_void_p), kl + ku + 1,
                                    x_np.ctypes.data_as(C.c_void_p), beta, y_np.ctypes.data_as(C.c_void_p))
        assert rc == 0
        want = oracle.gbmv("N", n, n, kl, ku, 1.5, A_np, x_np, beta, y0)
        assert np.max(np.abs(y_np - want)) <= tol_rows(np.float64, kl + ku + 1, 1.0, 1.0, 1.5, beta, 1.0)


def test_gbmv_large_linearity_and_sampled_rows(L, oracle):
    """At a size the oracle cannot sweep quickly: sampled rows vs the oracle's definition
    (regenerating only the needed entries from the counter-based stream) + linearity."""
    n, kl, ku = 1 << 24, 8, 8
    nb = kl + ku + 1
    A = L.brand(n, n, kl, ku, seed=0x1)
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    L.fill_uniform(x, 0x9)
    y = A @ x
    rows = np.unique(np.concatenate([np.arange(0, 40), np.arange(n - 40, n),
                                     np.random.default_rng(0).integers(0, n, 2000)]))
    yg = y[torch.from_numpy(rows).cuda()].cpu().numpy()
    for i, got in zip(rows.tolist(), yg.tolist()):
        jlo, jhi = max(0, i - kl), min(n - 1, i + ku)
        acc = 0.0
        for j in range(jlo, jhi + 1):
            a = oracle.fill_uniform(0x1, 1, offset=(ku + i - j) + nb * j)[0]
            acc += a * oracle.fill_uniform(0x9, 1, offset=j)[0]
        assert abs(got - acc) <= tol_rows(np.float64, nb, 1.0, 1.0), i
    # linearity: A(2x) == 2 A x exactly (power-of-two scaling)
    y2 = A @ (2.0 * x)
    assert torch.equal(y2, 2.0 * y)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_gbmv_multi_banded_times_dense(L, oracle, dt):
    """`mul!(C, A::BandedMatrix, B::Matrix, alpha, beta)` -> lbm_?gbmv_multi: groups of 8/4/2 columns share one
    pass over A (narrow and wide bands, both ops: wide op T through gbmv_wide_t_multi_kernel since round 2); both ops, ragged nrhs, padded ldx/ldy, and NaNs
    parked in A's out-of-matrix slots."""
    npdt = NP[dt]
    cases = [  # (m, n, kl, ku, nrhs, pad)
        (1000, 1000, 1, 1, 8, 0), (3001, 2999, 4, 2, 11, 4), (777, 777, 8, 8, 3, 0), (5000, 5003, 16, 16, 5, 8),
        (40, 33, 5, 2, 2, 0), (9, 9, 3, 0, 7, 0), (2000, 2000, 128, 128, 3, 0), (70001, 70001, 2, 2, 13, 0), (5, 5, 1, 1, 1, 0),
        (300, 300, 0, 0, 4, 0),
        # wide bands: gbmv_wide_n_multi_kernel / gbmv_wide_t_multi_kernel (one pass over A per 8 / 4 / 2 columns)
        (3000, 2900, 40, 30, 11, 4), (4097, 4097, 128, 128, 8, 0), (1500, 1700, 100, 3, 2, 0), (2500, 2500, 17, 17, 6, 2),
    ]
    for s, (m, n, kl, ku, nrhs, pad) in enumerate(cases):
        lda = kl + ku + 1
        A_np = np.asfortranarray(oracle.fill_uniform(200 + s, lda * n, npdt).reshape((lda, n), order="F"))
        Ap = A_np.copy()
        r = np.arange(lda)[:, None]
        i = np.arange(n)[None, :] + r - ku
        Ap[(i < 0) | (i >= m)] = np.nan
        A = L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(Ap.T)).cuda(), m, kl, ku)
        for trans in ("N", "T"):
            nin, nout = (n, m) if trans == "N" else (m, n)
            ldx, ldy = nin + pad, nout + pad
            X_np = oracle.fill_uniform(300 + s, ldx * nrhs, npdt).reshape(nrhs, ldx)     # row c = column c of X
            Y_np = oracle.fill_uniform(400 + s, ldy * nrhs, npdt).reshape(nrhs, ldy)
            for alpha, beta in ((1.0, 0.0), (-1.5, 0.5)):
                Xd = torch.from_numpy(X_np).cuda()
                Yd = torch.from_numpy(Y_np.copy()).cuda()
                # column-major views with padded leading dimensions: no copies inside the host layer
                L.mul_(Yd[:, :nout].t(), A if trans == "N" else A.T, Xd[:, :nin].t(), alpha, beta)
                got = Yd.cpu().numpy()
                tol = tol_rows(npdt, lda, 1.0, 1.0, alpha, beta, 1.0)
                for c in range(nrhs):
                    want = oracle.gbmv(trans, m, n, kl, ku, alpha, A_np, X_np[c, :nin].copy(), beta, Y_np[c, :nout].copy())
                    assert np.max(np.abs(got[c, :nout] - want)) <= tol, (m, n, kl, ku, nrhs, trans, c)
                    if pad:
                        assert np.array_equal(got[c, nout:], Y_np[c, nout:])           # padding rows untouched
    with pytest.raises(L.DimensionMismatch):
        L.mul_(torch.empty(5, 3, dtype=dt, device="cuda"), L.brand(5, 5, 1, 1, dtype=dt), torch.empty(5, 4, dtype=dt, device="cuda"))
