"""Parity at BASELINE.json's FULL sizes.  The oracle cannot sweep these in seconds, so each config is
checked through (a) entries sampled at random plus both ends, recomputed by the oracle from the
counter-based input stream (only the needed inputs are regenerated), and (b) size-independent
properties: exact linearity under power-of-two scaling, and a column-window re-computation for
the banded x banded product."""
import numpy as np
import pytest
import torch

from tests._util import tol_rows

pytestmark = pytest.mark.gpu


def _rows(n, k=1500, ends=64, seed=0):
    return np.unique(np.concatenate([np.arange(0, ends), np.arange(n - ends, n),
                                     np.random.default_rng(seed).integers(0, n, k)]))


@pytest.mark.parametrize("lu", [1, 8])
def test_metric_gbmv_n2_27(L, oracle, lu):
    """Metric config: A = brand(n,n,l,u) Float64, n = 2^27, l = u in {1, 8}."""
    n, kl, ku = 1 << 27, lu, lu
    nb = kl + ku + 1
    A = L.brand(n, n, kl, ku, seed=0x1)
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    L.fill_uniform(x, 0x9)
    y = A @ x
    rows = _rows(n)
    yg = y[torch.from_numpy(rows).cuda()].cpu().numpy()
    for i, got in zip(rows.tolist(), yg.tolist()):
        js = np.arange(max(0, i - kl), min(n - 1, i + ku) + 1)
        a = np.array([oracle.fill_uniform_numpy(0x1, 1, np.float64, int((ku + i - j) + nb * j))[0] for j in js])
        xv = oracle.fill_uniform_numpy(0x9, len(js), np.float64, int(js[0]))
        acc = 0.0
        for aa, xx in zip(a, xv):
            acc += aa * xx
        assert abs(got - acc) <= tol_rows(np.float64, nb, 1.0, 1.0), i
    assert torch.equal(A @ (4.0 * x), 4.0 * y)            # exact linearity (power-of-two scaling)
    ones = torch.ones_like(x)                               # adjoint identity: <A^T x, 1> == <x, A 1>
    s1 = torch.dot(A.T @ x, ones).item()
    s2 = torch.dot(x, A @ ones).item()
    assert abs(s1 - s2) <= 1e-9 * abs(s2)


def test_config2_symtridiagonal_bidiagonal_n2_28(L, oracle):
    n = 1 << 28
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    e = torch.empty(n - 1, dtype=torch.float64, device="cuda")
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    for t, s in ((d, 1), (e, 2), (x, 4)):
        L.fill_uniform(t, s)
    rows = _rows(n)
    idx = torch.from_numpy(rows).cuda()

    def val(seed, i):
        return oracle.fill_uniform_numpy(seed, 1, np.float64, int(i))[0]

    for kind, Aop in (("sym", L.SymTridiagonal(d, e)), ("U", L.Bidiagonal(d, e, "U")), ("L", L.Bidiagonal(d, e, "L"))):
        y = Aop @ x
        yg = y[idx].cpu().numpy()
        for i, got in zip(rows.tolist(), yg.tolist()):
            acc = 0.0
            if kind in ("sym", "L") and i > 0:
                acc += val(2, i - 1) * val(4, i - 1)
            acc += val(1, i) * val(4, i)
            if kind in ("sym", "U") and i < n - 1:
                acc += val(2, i) * val(4, i + 1)
            assert abs(got - acc) <= tol_rows(np.float64, 3, 1.0, 1.0), (kind, i)
        assert torch.equal(Aop @ (2.0 * x), 2.0 * y)
        del y


def test_config3_laplacian_8192(L, oracle):
    g = 8192
    h = 1.0 / g
    D = L.banded_from_diagonals(g, {0: -2.0 / h**2, 1: 1.0 / h**2, -1: 1.0 / h**2})
    x = torch.empty(g * g, dtype=torch.float64, device="cuda")
    L.fill_uniform(x, 5)
    Lap = L.KronSum.from_broadcast(L.BlockKron(D, L.Eye(g)), L.BlockKron(L.Eye(g), D))
    y = Lap @ x
    rng = np.random.default_rng(1)
    pts = [(0, 0), (g - 1, g - 1), (0, g - 1), (g - 1, 0)] + [tuple(p) for p in rng.integers(0, g, (1500, 2))]
    flat = torch.tensor([r + g * c for r, c in pts], dtype=torch.int64, device="cuda")
    yg = y[flat].cpu().numpy()
    c2 = 1.0 / h**2

    def X(r, c):
        return oracle.fill_uniform_numpy(5, 1, np.float64, int(r + g * c))[0] if 0 <= r < g and 0 <= c < g else 0.0

    for (r, c), got in zip(pts, yg.tolist()):
        # oracle order: A-term q = c-1, c, c+1 then B-term p = r-1, r, r+1 (oracle/lbm_oracle_impl.h kronsum2d_mv)
        acc = 0.0
        for q, co in ((c - 1, c2), (c, -2 * c2), (c + 1, c2)):
            if 0 <= q < g:
                acc += co * X(r, q)
        for p, co in ((r - 1, c2), (r, -2 * c2), (r + 1, c2)):
            if 0 <= p < g:
                acc += co * X(p, c)
        assert abs(got - acc) <= tol_rows(np.float64, 6, 2 * c2, 1.0), (r, c)
    assert torch.equal(Lap @ (2.0 * x), 2.0 * y)


def test_config4_gbmm_n2_20_wide(L, oracle):
    """C = A*B, l = u = 128, n = 2^20 (DMMA path): column windows recomputed by the oracle on the
    sub-product that the column-partition logic defines (lbm_b200.shard.GbmmColumnShard)."""
    from lbm_b200.shard import GbmmColumnShard
    n, l = 1 << 20, 128
    A = L.brand(n, n, l, l, seed=1)
    B = L.brand(n, n, l, l, seed=2)
    C = A @ B
    assert C.bandwidths() == (256, 256)
    lda = 2 * l + 1
    for c0 in (0, 4096 * 77, n - 64):
        p = GbmmColumnShard(n, l, l, l, l, 0, 1, c0, c0 + 64)
        (kla, kua), (klb, kub), (klc, kuc) = p.sub_bandwidths()
        As = np.asfortranarray(oracle.fill_uniform(1, lda * (p.pb - p.pa), offset=lda * p.pa).reshape((lda, p.pb - p.pa), order="F"))
        Bs = np.asfortranarray(oracle.fill_uniform(2, lda * 64, offset=lda * c0).reshape((lda, 64), order="F"))
        want = oracle.gbmm(p.rb - p.ra, 64, p.pb - p.pa, As, kla, kua, Bs, klb, kub, klc=klc, kuc=kuc)
        got = np.asfortranarray(C.data[c0:c0 + 64].cpu().numpy().T)
        m_loc = p.rb - p.ra
        dg, dw = oracle.band_to_dense(got, m_loc, klc, kuc), oracle.band_to_dense(want, m_loc, klc, kuc)
        assert np.max(np.abs(dg - dw)) <= tol_rows(np.float64, lda, 1.0, 1.0), c0


def test_config5_batched_chain_4096(L, oracle):
    """4096 independent A*B*C, brand(4096,4096,4,4) Float32: sampled batch items vs the oracle."""
    batch, n = 4096, 4096
    fs = [torch.empty((batch, n, 9), dtype=torch.float32, device="cuda") for _ in range(3)]
    for i, f in enumerate(fs):
        L.fill_uniform(f, 1 + i)
    D = L.chain3_batched(*fs, [4, 4, 4], [4, 4, 4])
    assert tuple(D.shape) == (batch, n, 25)
    tol = tol_rows(np.float32, 9, 1.0, 1.0) * 9 + tol_rows(np.float32, 9, 9.0, 1.0)
    for b in (0, 1, 777, 2048, 4095):
        items = [oracle.fill_uniform(1 + i, n * 9, np.float32, offset=b * n * 9).reshape(1, n, 9) for i in range(3)]
        want = oracle.gbmm_chain3_batched(n, [4, 4, 4], [4, 4, 4], *items)[0]
        got = D[b].cpu().numpy()
        dg = oracle.band_to_dense(np.asfortranarray(got.T), n, 12, 12)
        dw = oracle.band_to_dense(np.asfortranarray(want.T), n, 12, 12)
        assert np.max(np.abs(dg - dw)) <= tol, b
