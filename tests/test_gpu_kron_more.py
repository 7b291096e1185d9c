"""Round-2 Kronecker coverage on the GPU: Float32 kron2d / kron3d, the 9-diagonal and one-sided band families on the
streaming ring kernel, batched mode products at sizes the ring kernel takes (n >= 512 rows), and the DiagTrav device
kernels (gather, inverse, zero_bottomright) against the reference's index map (oracle.diagtrav)."""
import numpy as np
import pytest
import torch

from tests._util import tol_rows

pytestmark = pytest.mark.gpu
NP = {torch.float64: np.float64, torch.float32: np.float32}


def _band(L, oracle, n, l, u, seed, dt):
    A_np = oracle.brand(n, n, l, u, seed=seed, dtype=NP[dt])
    return A_np, L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(A_np.T)).cuda(), n, l, u)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("case", [(600, 520, (1, 1), (1, 1)), (520, 1024, (2, 2), (2, 2)), (300, 512, (4, 4), (4, 4)),
                                  (96, 80, (1, 1), (1, 1)), (400, 768, (2, 2), (1, 1)), (257, 516, (1, 0), (0, 2)), (130, 512, (4, 4), (0, 0))])
def test_kron2d_both_eltypes_and_band_families(L, oracle, dt, case):
    """kron(A,B) x = vec(B X A^T) (BlockKron(A,B)*x; the arithmetic of KronTrav(A,B)*DiagTrav(X), src/blockkron.jl:433-440)."""
    n1, n2, (kla, kua), (klb, kub) = case
    A_np, A = _band(L, oracle, n1, kla, kua, 31, dt)
    B_np, B = _band(L, oracle, n2, klb, kub, 32, dt)
    x = oracle.fill_uniform(5, n1 * n2, NP[dt])
    want = oracle.kron2d_mv(n1, n2, A_np, kla, kua, B_np, klb, kub, x)
    got = (L.BlockKron(A, B) @ torch.from_numpy(x).cuda()).cpu().numpy()
    tol = tol_rows(NP[dt], (kla + kua + 1) * (klb + kub + 1), 1.0, 1.0) * 2
    assert np.max(np.abs(got - want)) <= tol


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("n,bw", [(8, (1, 1)), (33, (2, 1)), (512, (1, 1)), (512, (2, 2))])
def test_kron3d_both_eltypes(L, oracle, dt, n, bw):
    """The three mode products of src/blockkron.jl:441-450; at n = 512 all three sweeps run on the streaming ring kernel
    (mode 2 batched over the slices) -- checked on sampled entries against a numpy restatement of the same sweeps."""
    facs = [_band(L, oracle, n, bw[0], bw[1], 40 + i, dt) for i in range(3)]
    rng = np.random.default_rng(2)
    x = torch.from_numpy(rng.random(n * n * n).astype(NP[dt])).cuda()
    y = torch.empty_like(x)
    fn = getattr(L.lib(), "lbm_dkron3d_mv" if dt == torch.float64 else "lbm_skron3d_mv")
    from lbm_b200.banded import _ft, _ptr
    h = L._lib.handle_of(x)
    A, B, Cm = (f[1] for f in facs)
    ft = _ft(dt)
    h.check(fn(h.ptr, n, A.l, A.u, _ptr(A.data), A.lda, B.l, B.u, _ptr(B.data), B.lda, Cm.l, Cm.u, _ptr(Cm.data), Cm.lda,
               ft(1.0), _ptr(x), ft(0.0), _ptr(y)), "kron3d")
    X = x.cpu().numpy().astype(np.float64).reshape(n, n, n)      # [l, j, k] (k fastest)
    Y = y.cpu().numpy().reshape(n, n, n)
    dA, dB, dC = (oracle.band_to_dense(f[0].astype(np.float64), n, *bw) for f in facs)
    pts = [(0, 0, 0), (n - 1, n - 1, n - 1), (n // 2, 1, n - 2)] + [tuple(p) for p in rng.integers(0, n, (200, 3))]
    nb = bw[0] + bw[1] + 1
    tol = tol_rows(NP[dt], nb ** 3, 1.0, 1.0) * 3
    for (l, j, k) in pts:
        # Y[l,j,k] = sum_{l',j',k'} A[l,l'] B[j,j'] C[k,k'] X[l',j',k'] restricted to the bands
        ls = range(max(0, l - bw[0]), min(n, l + bw[1] + 1))
        js = range(max(0, j - bw[0]), min(n, j + bw[1] + 1))
        ks = range(max(0, k - bw[0]), min(n, k + bw[1] + 1))
        sub = X[np.ix_(list(ls), list(js), list(ks))]
        want = np.einsum("a,b,c,abc->", dA[l, list(ls)], dB[j, list(js)], dC[k, list(ks)], sub)
        assert abs(Y[l, j, k] - want) <= tol, (l, j, k)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_diagtrav_device_kernels(L, oracle, dt):
    """lbm_?diagtrav / lbm_?invdiagtrav / lbm_?zero_bottomright == the reference's traversal (oracle.diagtrav, pinned by the
    golden vectors of test/test_blockkron.jl:26-35, :53-57, :64-69), bit for bit (pure data movement)."""
    rng = np.random.default_rng(3)
    for shape in [(3, 3), (2, 3), (3, 2), (1, 5), (5, 1), (4, 7), (130, 77), (1000, 1000), (4, 4, 4), (7, 7, 7), (40, 40, 40)]:
        X = rng.standard_normal(shape).astype(NP[dt])
        D = L.DiagTrav(torch.from_numpy(X).cuda())
        want = oracle.diagtrav(X)
        got = D.vector().cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want), shape
        assert len(D) == len(want)
        # the constructor zeroed the bottom right (zero_bottomright, src/blockkron.jl:27-64)
        assert np.array_equal(D.array.cpu().numpy(), oracle.zero_bottomright(X)), shape
        # inverse: InvDiagTrav(v) reproduces the zeroed array, and round-trips
        back = L.DiagTrav.from_vector(torch.from_numpy(want).cuda(), shape)
        assert np.array_equal(back.array.cpu().numpy(), oracle.zero_bottomright(X)), shape
        assert np.array_equal(back.vector().cpu().numpy(), want)
    with pytest.raises(L.DimensionMismatch):
        L.DiagTrav.from_vector(torch.zeros(5, dtype=dt, device="cuda"), (3, 3))


def test_krontrav_apply_float32(L, oracle):
    """K*DiagTrav(X) = DiagTrav(B*X*A') in Float32 (the reference sweeps element types, test/test_tridiag.jl:14)."""
    n = 64
    A_np, A = _band(L, oracle, n, 1, 1, 7, torch.float32)
    B_np, B = _band(L, oracle, n, 2, 1, 8, torch.float32)
    X = np.random.default_rng(4).random((n, n)).astype(np.float32)
    out = L.KronTrav(A, B).mul_diagtrav(L.DiagTrav(torch.from_numpy(X).cuda())).vector().cpu().numpy()
    dA, dB = oracle.band_to_dense(A_np, n, 1, 1).astype(np.float64), oracle.band_to_dense(B_np, n, 2, 1).astype(np.float64)
    want = oracle.diagtrav(dB @ oracle.zero_bottomright(X).astype(np.float64) @ dA.T)
    assert np.max(np.abs(out - want)) <= tol_rows(np.float32, 3 * 4, 1.0, 1.0) * 2
