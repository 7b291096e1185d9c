"""GPU parity: `+`, `-`, unary `-`, scalar `*`, UniformScaling and products among Diagonal / Bidiagonal /
Tridiagonal / SymTridiagonal (src/special.jl:75-222, src/tridiag.jl:206-211,:570-575, src/bidiag.jl:330-351),
through lbm_?tridiag_axpby / lbm_?tridiag_pack / lbm_?gbmm, against the oracle.  Sums are bit-exact."""
import itertools

import numpy as np
import pytest
import torch

from tests._util import tol_rows
from tests.test_special import KINDS, RETURNS, make

pytestmark = pytest.mark.gpu
NP = {torch.float64: np.float64, torch.float32: np.float32}


def _t(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def dev(L, X):
    k, dl, d, du = X
    if k == "D":
        return L.Diagonal(_t(d))
    if k == "BU":
        return L.Bidiagonal(_t(d), _t(du), "U")
    if k == "BL":
        return L.Bidiagonal(_t(d), _t(dl), "L")
    if k == "S":
        return L.SymTridiagonal(_t(d), _t(dl))
    return L.Tridiagonal(_t(dl), _t(d), _t(du))


def kind_of(L, R):
    from lbm_b200.special import _kind
    return _kind(R)


def host(R):
    dl, d, du = R._diagonals()
    m = max(R.n - 1, 0)
    c = lambda v: None if v is None else v[:m].cpu().numpy()
    return c(dl), d.cpu().numpy(), c(du)


def same(R, want):
    got = host(R)
    for g, w in zip(got, want[1:]):
        if w is None or g is None:
            assert (w is None or len(w) == 0) and (g is None or len(g) == 0)
            continue
        assert np.array_equal(g, w)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("n", [1, 2, 3, 33, 1000, 4097, (1 << 20) + 3])
def test_all_pairs_add_sub(L, oracle, dt, n):
    for ka, kb in itertools.product(KINDS, KINDS):
        A, B = make(oracle, ka, n, 11, NP[dt]), make(oracle, kb, n, 23, NP[dt])
        dA, dB = dev(L, A), dev(L, B)
        for sb, R in ((1.0, dA + dB), (-1.0, dA - dB)):
            want = oracle.special_axpby(1.0, A, sb, B)
            assert kind_of(L, R) == RETURNS[(ka, kb)] == want[0]
            if n == 1:
                assert np.array_equal(R.d.cpu().numpy(), want[2])
            else:
                same(R, want)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_uniformscaling_neg_scale(L, oracle, dt):
    n = 1025
    for k in KINDS:
        A = make(oracle, k, n, 5, NP[dt])
        dA = dev(L, A)
        for lam in (3.0, -0.625):
            I = L.UniformScaling(lam)
            same(dA + I, oracle.special_axpby(1.0, A, 1.0, ("I", lam)))
            same(I + dA, oracle.special_axpby(1.0, ("I", lam), 1.0, A))
            same(dA - I, oracle.special_axpby(1.0, A, -1.0, ("I", lam)))
            same(I - dA, oracle.special_axpby(1.0, ("I", lam), -1.0, A))
        neg = -dA
        same(neg, (k,) + tuple(None if v is None else -np.asarray(v) for v in A[1:]))
        sc = dA * 2.5
        same(sc, (k,) + tuple(None if v is None else np.asarray(v) * NP[dt](2.5) for v in A[1:]))
        # operands are not modified, and an untouched diagonal is shared like the reference's `A.ev`
        same(dA, A)
    Bu = dev(L, make(oracle, "BU", n, 5, NP[dt]))
    D = dev(L, make(oracle, "D", n, 7, NP[dt]))
    assert (Bu + D).ev.data_ptr() == Bu.ev.data_ptr()     # special.jl:75-78


def test_sym_ev_length_n_mismatch_throws(L, oracle):
    n = 9
    S = L.SymTridiagonal(_t(np.ones(n)), _t(np.ones(n)))           # len(ev) == n is a legal SymTridiagonal
    T = dev(L, make(oracle, "T", n, 3))
    with pytest.raises(L.DimensionMismatch):
        T + S                                                     # Julia: A.dl + B.ev broadcasts 8 against 9


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_pack_and_structured_products(L, oracle, dt):
    for n in (1, 2, 5, 64, 1000):
        for ka, kb in itertools.product(KINDS, KINDS):
            A, B = make(oracle, ka, n, 31, NP[dt]), make(oracle, kb, n, 37, NP[dt])
            dA, dB = dev(L, A), dev(L, B)
            PA = L.to_banded(dA, 2, 3)
            assert np.array_equal(PA.to_dense().cpu().numpy(), oracle.special_dense(A))
            C = dA @ dB
            la, ua = dA.bandwidths()
            lb, ub = dB.bandwidths()
            assert C.bandwidths() == (la + lb, ua + ub)
            want = oracle.special_dense(A).astype(np.float64) @ oracle.special_dense(B).astype(np.float64)
            err = np.max(np.abs(C.to_dense().cpu().numpy() - want)) if n else 0.0
            assert err <= tol_rows(NP[dt], 3, 0.5, 0.5), (ka, kb, n, err)   # test/test_special.jl:154-160


def test_sum_then_apply_matches_broadcast(L, oracle):
    # (A + B)*x through the materialised sum equals A*x + B*x within the stated tolerance
    n = 100003
    A, B = make(oracle, "T", n, 1), make(oracle, "BL", n, 2)
    x = oracle.fill_uniform(9, n)
    R = dev(L, A) + dev(L, B)
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    L.mul_(y, R, _t(x))
    want = oracle.tridiag_mv(n, A[1], A[2], A[3], x) + oracle.tridiag_mv(n, B[1], B[2], None, x)
    assert np.max(np.abs(y.cpu().numpy() - want)) <= tol_rows(np.float64, 5, 1.0, 1.0)


def test_special_abi_argument_errors(L):
    """LAPACK-style negative returns from the C ABI (never a throw / abort across the boundary)."""
    import ctypes as C
    lib, h = L.lib(), L.handle_for(0)
    d = torch.ones(8, dtype=torch.float64, device="cuda")
    e = torch.ones(7, dtype=torch.float64, device="cuda")
    out = torch.empty((8, 5), dtype=torch.float64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    assert lib.lbm_dtridiag_axpby(h.ptr, -1, 1.0, None, p(d), None, 1.0, None, p(d), None, 0.0, None, p(d), None) == -2
    assert lib.lbm_dtridiag_axpby(h.ptr, 8, 1.0, None, p(d), None, 1.0, None, p(d), None, 0.0, None, None, None) == -13
    assert lib.lbm_dtridiag_mm(h.ptr, 8, p(e), p(d), p(e), p(e), p(d), p(e), p(out), 4) == -10      # ldc < 5
    assert lib.lbm_dtridiag_mm(h.ptr, 8, p(e), None, p(e), p(e), p(d), p(e), p(out), 5) == -4
    assert lib.lbm_dtridiag_pack(h.ptr, 8, p(e), p(d), None, 0, 1, p(out), 5) == -6                 # subdiagonal but kl == 0
    assert lib.lbm_dtridiag_pack(h.ptr, 8, None, p(d), None, 1, 1, p(out), 2) == -9                 # lda < kl+ku+1
    i64 = C.c_int64 * 1
    assert lib.lbm_dgbmv_chain(h.ptr, 0, i64(8), i64(8), i64(0), i64(0), (C.c_void_p * 1)(d.data_ptr()), i64(1), 1.0, p(d), 0.0, p(d)) == -2
    assert "nf < 1" in lib.lbm_last_error_string(h.ptr).decode()
