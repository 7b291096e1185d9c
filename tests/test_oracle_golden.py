"""The oracle against every golden vector / KAT the reference's own tests hold for the path
(tests/golden/reference_kats.json, transcribed from /root/reference/test by make_reference_kats.py)."""
import numpy as np
import pytest


def _by_kind(kats, kind):
    out = [k for k in kats if k["kind"] == kind]
    assert out, kind
    return out


def test_diagtrav(kats, oracle):
    for k in _by_kind(kats, "diagtrav"):
        got = oracle.diagtrav(np.array(k["X"]))
        assert got.tolist() == k["expect"], k["src"]


def test_diagtrav_padded(kats, oracle):
    for k in _by_kind(kats, "diagtrav_padded"):
        C = np.zeros(k["shape"], dtype=np.int64)
        blk = np.array(k["block"])
        C[: blk.shape[0], : blk.shape[1]] = blk
        got = oracle.diagtrav(oracle.zero_bottomright(C))
        assert got.tolist() == k["expect"], k["src"]


def test_diagtrav3_block_order(kats, oracle):
    (k,) = _by_kind(kats, "diagtrav3_blocks")
    n = k["n"]
    X = np.arange(n**3).reshape(n, n, n)
    v = oracle.diagtrav(X).tolist()
    b2 = [X[a - 1, b - 1, c - 1] for a, b, c in k["block2"]]
    b3 = [X[a - 1, b - 1, c - 1] for a, b, c in k["block3"]]
    assert v[0] == X[0, 0, 0]
    assert v[1:4] == b2
    assert v[4:10] == b3


def test_invdiagtrav(kats, oracle):
    (k,) = _by_kind(kats, "invdiagtrav")
    X = np.array(k["X"])
    got = oracle.invdiagtrav(oracle.diagtrav(X), *X.shape)
    # DiagTrav keeps the upper-left triangle only (zero_bottomright)
    got = oracle.invdiagtrav(oracle.diagtrav(oracle.zero_bottomright(X.copy())), *X.shape)
    assert got.tolist() == k["expect"]


def test_krontrav_vec(kats, oracle):
    for k in _by_kind(kats, "krontrav_vec"):
        assert oracle.krontrav_vec(np.array(k["a"]), np.array(k["b"])).tolist() == k["expect"], k["src"]


def test_krontrav_mul(kats, oracle):
    (k,) = _by_kind(kats, "krontrav_mul")
    A, B, X, Y = (np.array(k[s]) for s in "ABXY")
    K = oracle.krontrav_matrix(A, B)
    lhs_x = K @ oracle.diagtrav(oracle.zero_bottomright(X.copy()))
    lhs_y = K @ oracle.diagtrav(oracle.zero_bottomright(Y.copy()))
    rhs = oracle.krontrav_mul_diagtrav(A, B, X)
    assert lhs_x.tolist() == lhs_y.tolist() == rhs.tolist()
    assert rhs.tolist() == [211, 291, 533]  # value re-derived in SURVEY.md 8(c)


def test_krontrav3_matrix(kats, oracle):
    (k,) = _by_kind(kats, "krontrav3_matrix")
    A, B, Cm = np.array(k["A"]), np.array(k["B"]), np.array(k["C"])
    n = 2
    idx = [(a, b, c) for a in range(n) for b in range(n) for c in range(n)]
    # columns of K from unit tensors inside the kept tetrahedron, in DiagTrav order
    order = oracle.diagtrav(np.arange(n**3).reshape(n, n, n))
    cols = []
    for flat in order:
        X = np.zeros((n, n, n))
        X[np.unravel_index(int(flat), (n, n, n))] = 1
        cols.append(oracle.krontrav3_mul_diagtrav(A, B, Cm, X))
    K = np.stack(cols, axis=1)
    assert K.astype(int).tolist() == k["expect"]
    assert len(idx) == 8


def _laplace(n, diag, off):
    D = np.zeros((n, n))
    np.fill_diagonal(D, diag)
    for i in range(n - 1):
        D[i, i + 1] = D[i + 1, i] = off
    return D


def test_krontrav_laplace2d(kats, oracle):
    rng = np.random.default_rng(0)
    for k in _by_kind(kats, "krontrav_laplace2d"):
        n = k["n"]
        D = _laplace(n, k["diag"], k["off"])
        I = np.eye(n)
        X = np.triu(rng.standard_normal((n, n)))[:, ::-1].copy()  # test/test_blockkron.jl:205
        if "X * Δ'" in k["snippet"]:
            got = oracle.krontrav_matrix(D, I) @ oracle.diagtrav(X)
            want = oracle.diagtrav(X @ D.T)
        else:
            got = oracle.krontrav_matrix(I, D) @ oracle.diagtrav(X)
            want = oracle.diagtrav(D @ X)
        assert np.array_equal(got, want), k["src"]


def test_kron_c_oracle_matches_identity(oracle):
    """C kron2d / kronsum2d restatements == dense numpy kron, on the +-1/-2 Laplacian (exact)
    and on random bands."""
    rng = np.random.default_rng(1)
    for (n1, n2, la, ua, lb, ub) in [(4, 4, 1, 1, 1, 1), (5, 3, 1, 1, 0, 0), (3, 6, 2, 1, 1, 2), (1, 1, 0, 0, 0, 0)]:
        A = rng.integers(-3, 4, size=(la + ua + 1, n1)).astype(np.float64, order="F")
        B = rng.integers(-3, 4, size=(lb + ub + 1, n2)).astype(np.float64, order="F")
        Ad, Bd = oracle.band_to_dense(A, n1, la, ua), oracle.band_to_dense(B, n2, lb, ub)
        x = rng.integers(-5, 6, size=n1 * n2).astype(np.float64)
        got = oracle.kron2d_mv(n1, n2, A, la, ua, B, lb, ub, x)
        assert np.array_equal(got, np.kron(Ad, Bd) @ x)
        got = oracle.kronsum2d_mv(n1, n2, A, la, ua, B, lb, ub, x)
        want = (np.kron(Ad, np.eye(n2)) + np.kron(np.eye(n1), Bd)) @ x
        assert np.array_equal(got, want)


def test_dense_structured(kats, oracle):
    for kind in ("dense_tridiagonal", "dense_symtridiagonal", "dense_bidiagonal"):
        for k in _by_kind(kats, kind):
            n = len(k["expect"])
            if kind == "dense_tridiagonal":
                dl, d, du = (np.array(k[s], dtype=float) for s in ("dl", "d", "du"))
            elif kind == "dense_symtridiagonal":
                d = np.array(k["dv"], dtype=float)
                dl = du = np.array(k["ev"], dtype=float)
            else:
                d = np.array(k["dv"], dtype=float)
                ev = np.array(k["ev"], dtype=float)
                dl, du = (None, ev) if k["uplo"] == "U" else (ev, None)
            dense = oracle.tridiag_mv(n, dl, d, du, np.eye(n))
            assert dense.tolist() == k["expect"], k["src"]


def test_symtridiag_kats(kats, oracle):
    (k,) = _by_kind(kats, "symtridiag_mv")
    d = np.array(k["dv"])
    ev = np.array(k["ev"])  # length n: only ev[:n-1] is used (src/tridiag.jl:11)
    got = oracle.tridiag_mv(1, None, d, None, np.array(k["x"]))
    assert got.tolist() == k["expect"]
    assert ev.shape[0] == 1
    (k,) = _by_kind(kats, "symtridiag_empty")
    got = oracle.tridiag_mv(0, None, np.zeros(0), None, np.ones((0, k["nrhs"])))
    assert got.shape == (0, 2)
    (k,) = _by_kind(kats, "symtridiag_cube")
    dv, ev = np.array(k["dv"], dtype=float), np.array(k["ev"], dtype=float)
    data, l, u = oracle.banded_from_diagonals(2, {0: dv, 1: ev, -1: ev})
    sq = oracle.gbmm(2, 2, 2, data, l, u, data, l, u)
    cube = oracle.gbmm(2, 2, 2, sq, 2, 2, data, l, u)
    assert oracle.band_to_dense(cube, 2, 3, 3).tolist() == k["expect"]


def test_sum_kats(kats, oracle):
    (k,) = _by_kind(kats, "tridiag_sum")
    n = len(k["d"])
    dense = oracle.tridiag_mv(n, np.array(k["dl"], float), np.array(k["d"], float), np.array(k["du"], float), np.eye(n))
    assert dense.sum() == k["expect"]
    (k,) = _by_kind(kats, "symtridiag_sum")
    ev = np.array(k["ev"], float)
    dense = oracle.tridiag_mv(n, ev, np.array(k["dv"], float), ev, np.eye(n))
    assert dense.sum() == k["expect"]


def test_sum_dims_restatement(kats, oracle):
    """oracle.tridiag_sum (restatement of Base._sum, src/tridiag.jl:594-660, src/bidiag.jl:396-435) reproduces the
    reference KATs `sum(T) == 24`, `sum(S) == 12` and `sum(T, dims=k) == sum(Matrix(T), dims=k)`
    (test/test_tridiag.jl:366-374), for all three kinds."""
    (k,) = _by_kind(kats, "tridiag_sum")
    dl, d, du = (np.array(k[f], float) for f in ("dl", "d", "du"))
    n = len(d)
    assert oracle.tridiag_sum(n, dl, d, du) == k["expect"]
    dense = oracle.tridiag_mv(n, dl, d, du, np.eye(n))
    assert np.array_equal(oracle.tridiag_sum(n, dl, d, du, 1), dense.sum(axis=0, keepdims=True))
    assert np.array_equal(oracle.tridiag_sum(n, dl, d, du, 2), dense.sum(axis=1, keepdims=True))
    (k,) = _by_kind(kats, "symtridiag_sum")
    dv, ev = np.array(k["dv"], float), np.array(k["ev"], float)
    assert oracle.tridiag_sum(n, ev, dv, ev) == k["expect"]
    rng = np.random.default_rng(5)
    for n in (1, 2, 7):
        d, e = rng.integers(-5, 6, n).astype(float), rng.integers(-5, 6, max(n - 1, 0)).astype(float)
        for (lo, up) in ((e, None), (None, e), (e, e)):           # Bidiagonal L, Bidiagonal U, SymTridiagonal
            dense = np.diag(d) + (np.diag(lo, -1) if lo is not None and n > 1 else 0) + (np.diag(up, 1) if up is not None and n > 1 else 0)
            assert oracle.tridiag_sum(n, lo, d, up) == dense.sum()
            assert np.array_equal(oracle.tridiag_sum(n, lo, d, up, 1), dense.sum(axis=0, keepdims=True))
            assert np.array_equal(oracle.tridiag_sum(n, lo, d, up, 2), dense.sum(axis=1, keepdims=True))


def test_krontrav_square(kats, oracle):
    (k,) = _by_kind(kats, "krontrav_square")
    n = k["n"]
    D = _laplace(n, k["diag"], k["off"])
    K = oracle.krontrav_matrix(D, np.eye(n))
    # (A^2) applied to DiagTrav(X) equals applying A twice via the B*X*A' identity
    rng = np.random.default_rng(3)
    X = np.triu(rng.integers(-4, 5, size=(n, n)).astype(float))[:, ::-1].copy()
    v = oracle.diagtrav(X)
    assert np.array_equal(K @ (K @ v), (K @ K) @ v)


def test_generator_c_vs_numpy(oracle):
    for dt in (np.float64, np.float32):
        a = oracle.fill_uniform(0x1, 4097, dt, offset=123)
        b = oracle.fill_uniform_numpy(0x1, 4097, dt, offset=123)
        assert np.array_equal(a, b)
        assert a.min() >= 0 and a.max() < 1
    assert abs(oracle.fill_uniform(7, 200000).mean() - 0.5) < 5e-3
