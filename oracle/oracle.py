"""CPU ORACLE (python face) -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does.

Two layers:

* thin ``ctypes`` wrappers over ``oracle/_build/liblbm_oracle.so`` (the C restatement
  in ``lbm_oracle.c`` / ``lbm_oracle_impl.h``; each C function cites the reference
  file:line it follows), and
* small numpy / pure-python restatements of the reference's *block-ordering* code
  (``DiagTrav``, ``InvDiagTrav``, ``KronTrav``), which are integer index maps and
  are pinned by the reference's golden vectors (``tests/golden/reference_kats.json``).

Parity status: see the header of ``lbm_oracle.c`` ("parity unpinned" for random-data
``gbmv``/``gbmm`` values; pinned by the reference's KATs elsewhere and by OpenBLAS).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblbm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc, a few seconds).  Returns the .so path."""
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("lbm_oracle.c", "lbm_oracle_impl.h", "Makefile")
    ):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_dtridiag_dot.restype = C.c_double
        _lib.oracle_stridiag_dot.restype = C.c_float
    return _lib


def _pfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "d"
    if dtype == np.float32:
        return "s"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _ct(dtype):
    return C.c_double if np.dtype(dtype) == np.float64 else C.c_float


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i(v):
    return C.c_int64(int(v))


# --------------------------------------------------------------------------------------
# generators
# --------------------------------------------------------------------------------------
def fill_uniform(seed: int, count: int, dtype=np.float64, offset: int = 0) -> np.ndarray:
    """The counter-based U[0,1) stream shared with the CUDA generator."""
    out = np.empty(int(count), dtype=dtype)
    fn = getattr(lib(), f"oracle_{_pfx(dtype)}fill_uniform")
    rc = fn(C.c_uint64(seed & (2**64 - 1)), _i(offset), _i(count), _p(out))
    assert rc == 0
    return out


def fill_uniform_numpy(seed: int, count: int, dtype=np.float64, offset: int = 0) -> np.ndarray:
    """Independent numpy restatement of the same stream (pins the C generator)."""
    with np.errstate(over="ignore"):
        idx = np.arange(offset, offset + count, dtype=np.uint64)
        z = np.uint64(seed & (2**64 - 1)) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    if np.dtype(dtype) == np.float64:
        return (z >> np.uint64(11)).astype(np.float64) * 2.0**-53
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0**-24)).astype(np.float32)


def stream_at(seed: int, idx, dtype=np.float64) -> np.ndarray:
    """The same counter-based stream evaluated at arbitrary (array of) 64-bit indices: what lets a full-size
    result (n = 2^27 ... 2^28) be checked row by row without regenerating the operands."""
    with np.errstate(over="ignore"):
        idx = np.asarray(idx).astype(np.uint64)
        z = np.uint64(seed & (2**64 - 1)) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    if np.dtype(dtype) == np.float64:
        return (z >> np.uint64(11)).astype(np.float64) * 2.0**-53
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0**-24)).astype(np.float32)


def gbmv_rows_stream(n: int, kl: int, ku: int, seed_a: int, seed_x: int, rows, x_scale: float = 1.0) -> np.ndarray:
    """(A x)[rows] for A = brand(n,n,kl,ku) (stream `seed_a`, tight lda) and x = x_scale * stream(`seed_x`): each row
    summed in ascending column order, as oracle_?gbmv does (lbm_oracle_impl.h), from regenerated entries only."""
    rows = np.asarray(rows, dtype=np.int64)
    nb = kl + ku + 1
    acc = np.zeros(rows.shape, dtype=np.float64)
    for q in range(nb):
        j = rows - kl + q
        ok = (j >= 0) & (j < n)
        jj = np.where(ok, j, 0)
        a = stream_at(seed_a, (kl + ku - q) + nb * jj)          # band row ku + i - j = kl + ku - q
        xv = stream_at(seed_x, jj) * x_scale
        acc = np.where(ok, acc + a * xv, acc)
    return acc


def tridiag_rows_stream(n: int, seed_dl, seed_d: int, seed_du, seed_x: int, rows, x_scale: float = 1.0) -> np.ndarray:
    """(A x)[rows] for Tridiagonal(dl, d, du) built from streams (None = that diagonal absent; SymTridiagonal passes
    the same seed twice): dl[i-1] x[i-1] + d[i] x[i] + du[i] x[i+1] in that order (src/tridiag.jl:471-484)."""
    rows = np.asarray(rows, dtype=np.int64)
    acc = np.zeros(rows.shape, dtype=np.float64)
    x = lambda i: stream_at(seed_x, i) * x_scale  # noqa: E731
    if seed_dl is not None:
        ok = rows > 0
        im = np.where(ok, rows - 1, 0)
        acc = np.where(ok, acc + stream_at(seed_dl, im) * x(im), acc)
    acc = acc + stream_at(seed_d, rows) * x(rows)
    if seed_du is not None:
        ok = rows < n - 1
        ip = np.where(ok, rows + 1, 0)
        acc = np.where(ok, acc + stream_at(seed_du, np.where(ok, rows, 0)) * x(ip), acc)
    return acc


def laplacian_points_stream(g: int, c2: float, seed_x: int, r, c, x_scale: float = 1.0) -> np.ndarray:
    """((kron(D,I)+kron(I,D)) x)[r + g*c] for D = c2*tridiag(1,-2,1), x = vec(X) from the stream: A-term columns
    c-1, c, c+1 then B-term rows r-1, r, r+1 (the order of oracle_?kronsum2d_mv)."""
    r = np.asarray(r, dtype=np.int64)
    c = np.asarray(c, dtype=np.int64)
    acc = np.zeros(r.shape, dtype=np.float64)

    def X(rr, cc):
        ok = (rr >= 0) & (rr < g) & (cc >= 0) & (cc < g)
        return np.where(ok, stream_at(seed_x, np.where(ok, rr + g * cc, 0)) * x_scale, 0.0), ok

    for (dr, dc, co) in ((0, -1, c2), (0, 0, -2 * c2), (0, 1, c2), (-1, 0, c2), (0, 0, -2 * c2), (1, 0, c2)):
        v, ok = X(r + dr, c + dc)
        acc = np.where(ok, acc + co * v, acc)
    return acc


def brand(m: int, n: int, l: int, u: int, seed: int, dtype=np.float64) -> np.ndarray:
    """`brand(m,n,l,u)` analogue: ALL (l+u+1)*n storage slots are U[0,1) ([up]
    BandedMatrices).  Returns the (l+u+1) x n column-major data
    array (Fortran order)."""
    flat = fill_uniform(seed, (l + u + 1) * n, dtype)
    return np.asfortranarray(flat.reshape((l + u + 1, n), order="F"))


# --------------------------------------------------------------------------------------
# dense <-> band helpers (numpy, for small cross-checks)
# --------------------------------------------------------------------------------------
def band_to_dense(data: np.ndarray, m: int, l: int, u: int) -> np.ndarray:
    """A[i,j] = data[u+i-j, j] for -u <= i-j <= l, zero elsewhere."""
    n = data.shape[1]
    A = np.zeros((m, n), dtype=data.dtype)
    for j in range(n):
        for i in range(max(0, j - u), min(m, j + l + 1)):
            A[i, j] = data[u + i - j, j]
    return A


def dense_to_band(A: np.ndarray, l: int, u: int, fill=0.0) -> np.ndarray:
    m, n = A.shape
    data = np.full((l + u + 1, n), fill, dtype=A.dtype, order="F")
    for j in range(n):
        for i in range(max(0, j - u), min(m, j + l + 1)):
            data[u + i - j, j] = A[i, j]
    return data


def banded_from_diagonals(n: int, diags: dict, dtype=np.float64):
    """`BandedMatrix(k1 => v1, k2 => v2, ...)` ([up] BandedMatrices; used at
    /root/reference/test/test_blockkron.jl:96,299).  Returns (data, l, u)."""
    ks = list(diags)
    l = max(0, -min(ks))
    u = max(0, max(ks))
    data = np.zeros((l + u + 1, n), dtype=dtype, order="F")
    for k, v in diags.items():
        v = np.broadcast_to(np.asarray(v, dtype=dtype), (n - abs(k),))
        if k >= 0:
            data[u - k, k:] = v
        else:
            data[u - k, : n + k] = v
    return data, l, u


# --------------------------------------------------------------------------------------
# C oracle wrappers
# --------------------------------------------------------------------------------------
def gbmv(trans, m, n, kl, ku, alpha, A, x, beta=0.0, y=None, lda=None):
    dt = A.dtype
    lda = A.shape[0] if lda is None else lda
    ylen = n if trans in ("T", "C", "t", "c") else m
    y = np.zeros(ylen, dtype=dt) if y is None else np.array(y, dtype=dt, copy=True)
    fn = getattr(lib(), f"oracle_{_pfx(dt)}gbmv")
    ct = _ct(dt)
    A = np.asfortranarray(A)
    x = np.ascontiguousarray(x, dtype=dt)
    rc = fn(C.c_char(trans.encode()), _i(m), _i(n), _i(kl), _i(ku), ct(alpha), _p(A), _i(lda),
            _p(x), ct(beta), _p(y))
    assert rc == 0, rc
    return y


def tridiag_mv(n, dl, d, du, X, alpha=1.0, beta=0.0, Y=None):
    dt = d.dtype
    X = np.asfortranarray(X, dtype=dt)
    X2 = X.reshape(n, -1, order="F") if X.ndim == 1 else X
    nrhs = X2.shape[1]
    Yo = np.zeros_like(X2, order="F") if Y is None else np.array(
        np.asarray(Y, dtype=dt).reshape(n, -1, order="F"), order="F", copy=True)
    fn = getattr(lib(), f"oracle_{_pfx(dt)}tridiag_mv")
    ct = _ct(dt)
    dl = None if dl is None else np.ascontiguousarray(dl, dtype=dt)
    du = None if du is None else np.ascontiguousarray(du, dtype=dt)
    d = np.ascontiguousarray(d, dtype=dt)
    rc = fn(_i(n), _p(dl), _p(d), _p(du), ct(alpha), _p(X2), _i(max(n, 1)), _i(nrhs), ct(beta),
            _p(Yo), _i(max(n, 1)))
    assert rc == 0, rc
    return Yo.reshape(X.shape, order="F")


def gbmm(m, n, k, A, kla, kua, B, klb, kub, klc=None, kuc=None, alpha=1.0, beta=0.0, Cin=None):
    dt = A.dtype
    klc = kla + klb if klc is None else klc
    kuc = kua + kub if kuc is None else kuc
    A = np.asfortranarray(A)
    B = np.asfortranarray(B)
    Cd = (np.zeros((klc + kuc + 1, n), dtype=dt, order="F") if Cin is None
          else np.array(Cin, dtype=dt, order="F", copy=True))
    fn = getattr(lib(), f"oracle_{_pfx(dt)}gbmm")
    ct = _ct(dt)
    rc = fn(_i(m), _i(n), _i(k), ct(alpha), _i(kla), _i(kua), _p(A), _i(A.shape[0]),
            _i(klb), _i(kub), _p(B), _i(B.shape[0]), ct(beta), _i(klc), _i(kuc), _p(Cd),
            _i(Cd.shape[0]))
    assert rc == 0, rc
    return Cd


def gbmm_chain3_batched(n, kl, ku, A, B, Cm):
    """A, B, Cm: arrays of shape (batch, n, kl_i+ku_i+1) C-contiguous == per item a
    column-major (kl+ku+1) x n band array.  Returns D (batch, n, sum+1)."""
    dt = A.dtype
    batch = A.shape[0]
    kl = np.asarray(kl, dtype=np.int64)
    ku = np.asarray(ku, dtype=np.int64)
    ldd = int(kl.sum() + ku.sum() + 1)
    D = np.zeros((batch, n, ldd), dtype=dt)
    fn = getattr(lib(), f"oracle_{_pfx(dt)}gbmm_chain3_batched")
    A = np.ascontiguousarray(A); B = np.ascontiguousarray(B); Cm = np.ascontiguousarray(Cm)
    rc = fn(_i(batch), _i(n), _p(kl), _p(ku), _p(A), _i(A[0].size), _p(B), _i(B[0].size),
            _p(Cm), _i(Cm[0].size), _p(D), _i(D[0].size))
    assert rc == 0, rc
    return D


def _kron_call(name, n1, n2, A, kla, kua, B, klb, kub, x, alpha, beta, y):
    dt = x.dtype
    A = np.asfortranarray(A, dtype=dt)
    B = np.asfortranarray(B, dtype=dt)
    x = np.ascontiguousarray(x)
    y = np.zeros(n1 * n2, dtype=dt) if y is None else np.array(y, dtype=dt, copy=True)
    fn = getattr(lib(), f"oracle_{_pfx(dt)}{name}")
    ct = _ct(dt)
    rc = fn(_i(n1), _i(n2), _i(kla), _i(kua), _p(A), _i(A.shape[0]), _i(klb), _i(kub), _p(B),
            _i(B.shape[0]), ct(alpha), _p(x), ct(beta), _p(y))
    assert rc == 0, rc
    return y


def kronsum2d_mv(n1, n2, A, kla, kua, B, klb, kub, x, alpha=1.0, beta=0.0, y=None):
    """(kron(A,I_n2) + kron(I_n1,B)) x = vec(X A^T + B X)."""
    return _kron_call("kronsum2d_mv", n1, n2, A, kla, kua, B, klb, kub, x, alpha, beta, y)


def kron2d_mv(n1, n2, A, kla, kua, B, klb, kub, x, alpha=1.0, beta=0.0, y=None):
    """kron(A,B) x = vec(B X A^T)."""
    return _kron_call("kron2d_mv", n1, n2, A, kla, kua, B, klb, kub, x, alpha, beta, y)


def tridiag_solve(dl, d, du, B):
    """`Tridiagonal(dl,d,du) \\ B`.  The reference has no solve code of its own: `\\` goes through ArrayLayouts to
    LinearAlgebra's `lu(::Tridiagonal)` = LAPACK ?gttrf/?gttrs (partial pivoting; exercised at
    /root/reference/test/test_tridiag.jl:345-363).  The oracle is therefore LAPACK ?gtsv itself (scipy), the same
    pivoted elimination; a missing diagonal (Bidiagonal) is zeros."""
    from scipy.linalg import lapack
    d = np.asarray(d)
    n = len(d)
    z = np.zeros(max(n - 1, 0), dtype=d.dtype)
    dl = z if dl is None else np.asarray(dl)[: max(n - 1, 0)]
    du = z if du is None else np.asarray(du)[: max(n - 1, 0)]
    B2 = np.asarray(B, dtype=d.dtype).reshape(n, -1)
    if n == 0:
        return np.asarray(B).copy()
    if n == 1:
        return (B2 / d[0]).reshape(np.asarray(B).shape)
    fn = lapack.dgtsv if d.dtype == np.float64 else lapack.sgtsv
    _, _, _, X, info = fn(dl.copy(), d.copy(), du.copy(), np.asfortranarray(B2.copy()))
    assert info == 0, info
    return X.reshape(np.asarray(B).shape)


def tridiag_dot(x, dl, d, du, y):
    dt = d.dtype
    fn = getattr(lib(), f"oracle_{_pfx(dt)}tridiag_dot")
    arrs = [np.ascontiguousarray(a, dtype=dt) for a in (x, dl, d, du, y)]
    return float(fn(_i(len(d)), *[_p(a) for a in arrs]))


def tridiag_sum(n, dl, d, du, dims=None):
    """Base._sum(A, dims) for Tridiagonal(dl,d,du) -- numpy restatement of /root/reference/src/tridiag.jl:594-628
    (SymTridiagonal: dl = du = ev[:n-1], :630-660; Bidiagonal: the missing diagonal is None, src/bidiag.jl:396-435).
    dims=None: scalar; dims=1: column sums (1 x n); dims=2: row sums (n x 1)."""
    d = np.asarray(d)
    z = np.zeros(max(n - 1, 0), dtype=d.dtype)
    dl = z if dl is None else np.asarray(dl)[: max(n - 1, 0)]
    du = z if du is None else np.asarray(du)[: max(n - 1, 0)]
    if dims is None:
        return d.sum() + dl.sum() + du.sum()
    res = d.copy()
    if n > 1:
        if dims == 1:      # res[j] = du[j-1] + d[j] + dl[j]
            res[1:] += du
            res[:-1] += dl
        else:              # res[i] = dl[i-1] + d[i] + du[i]
            res[1:] += dl
            res[:-1] += du
    return res.reshape(1, n) if dims == 1 else res.reshape(n, 1)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# --------------------------------------------------------------------------------------
# DiagTrav / InvDiagTrav / KronTrav (numpy restatements of the reference's index maps)
# --------------------------------------------------------------------------------------
def diagtrav(X: np.ndarray) -> np.ndarray:
    """`DiagTrav(X)`: traverse anti-diagonals (/root/reference/src/blockkron.jl:14-24;
    block K of a matrix = X[K:-1:1, 1:K] clipped to size, :101-133; axes :68-80).
    Golden: DiagTrav([1 2 3;4 5 6;7 8 9]) == [1,4,2,7,5,3] (test/test_blockkron.jl:25)."""
    X = np.asarray(X)
    if X.ndim == 2:
        m, n = X.shape
        out = []
        for K in range(1, max(m, n) + 1):
            # entries (k, j) with k + j == K + 1 (1-based), k descending, inside the matrix
            for j in range(1, K + 1):
                k = K + 1 - j
                if k <= m and j <= n:
                    out.append(X[k - 1, j - 1])
        return np.array(out, dtype=X.dtype)
    if X.ndim == 3:
        m, n, p = X.shape
        assert m == n == p
        out = []
        for K in range(1, n + 1):
            # block K: (k,j,l) with k+j+l == K+2; order per test/test_blockkron.jl:53-56:
            # k descending outermost, then j descending (l ascending).
            for k in range(K, 0, -1):
                for j in range(K + 1 - k, 0, -1):
                    l = K + 2 - k - j
                    out.append(X[k - 1, j - 1, l - 1])
        return np.array(out, dtype=X.dtype)
    raise ValueError("DiagTrav supports 2-D and 3-D arrays")


def invdiagtrav(v: np.ndarray, m: int, n: int) -> np.ndarray:
    """`invdiagtrav`: inverse of diagtrav onto an m x n matrix, zero bottom-right
    (golden: test/test_blockkron.jl:89 -> [1 2 3;4 5 0;7 0 0])."""
    X = np.zeros((m, n), dtype=np.asarray(v).dtype)
    it = iter(np.asarray(v).tolist())
    for K in range(1, max(m, n) + 1):
        for j in range(1, K + 1):
            k = K + 1 - j
            if k <= m and j <= n:
                try:
                    X[k - 1, j - 1] = next(it)
                except StopIteration:
                    return X
    return X


def krontrav_vec(a, b) -> np.ndarray:
    """`KronTrav(a,b)` for vectors == DiagTrav(b*a') (test/test_blockkron.jl:121-123)."""
    return diagtrav(np.outer(b, a))


def krontrav_matrix(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Dense `KronTrav(A,B)`: the matrix K with K*diagtrav(X) == diagtrav(B*X*A') for
    every X supported on the upper-left triangle (src/blockkron.jl:272-278, :433-440).
    Built column by column from that defining identity, restricted to in-triangle
    entries (block (K,J) = A[1:K,1:J] .* rot180(B[1:K,1:J]) pattern, :323-331)."""
    nA, mA = A.shape
    nB, mB = B.shape
    # input ordering: DiagTrav of an (mB x mA) matrix; output: (nB x nA)
    cols_idx = _diagtrav_index(mB, mA)
    rows_idx = _diagtrav_index(nB, nA)
    K = np.zeros((len(rows_idx), len(cols_idx)), dtype=np.result_type(A, B))
    for c, (p, q) in enumerate(cols_idx):       # X = e_p e_q^T
        # (B X A')[r, s] = B[r,p] * A[s,q]
        for r_, (r, s) in enumerate(rows_idx):
            K[r_, c] = B[r, p] * A[s, q]
    return K


def _diagtrav_index(m: int, n: int):
    """0-based (row, col) pairs in DiagTrav order, restricted to the triangle
    k + j <= max(m,n) + 1 that DiagTrav keeps (zero_bottomright, src/blockkron.jl:27-47)."""
    idx = []
    for K in range(1, max(m, n) + 1):
        for j in range(1, K + 1):
            k = K + 1 - j
            if k <= m and j <= n:
                idx.append((k - 1, j - 1))
    return idx


def krontrav_mul_diagtrav(A: np.ndarray, B: np.ndarray, X: np.ndarray) -> np.ndarray:
    """`KronTrav(A,B) * DiagTrav(X) = DiagTrav(B*X*A')` (src/blockkron.jl:433-440);
    DiagTrav zeroes X's bottom-right first (src/blockkron.jl:27-47)."""
    X = zero_bottomright(np.array(X, copy=True))
    return diagtrav(B @ X @ A.T)


def zero_bottomright(X: np.ndarray) -> np.ndarray:
    """src/blockkron.jl:38-47: zero entries with k + j > max(m,n) + 1 (1-based); 3-D cubic arrays (:57-64): zero
    X[k+1, j+1, l+1] for k >= n - (l + j) (0-based k, j, l).  In place, like `zero_bottomright!`."""
    if X.ndim == 3:
        n = X.shape[0]
        assert X.shape == (n, n, n)
        for ell in range(n):
            for j in range(n):
                for k in range(max(0, n - (ell + j)), n):
                    X[k, j, ell] = 0
        return X
    m, n = X.shape
    mu = max(m, n)
    for j in range(1, n + 1):
        for k in range(mu - j + 2, m + 1):
            X[k - 1, j - 1] = 0
    return X


def krontrav3_mul_diagtrav(A, B, Cm, X):
    """3-D mode products of src/blockkron.jl:441-450: A on dim 3, B on dim 2, C on dim 1."""
    n = X.shape[0]
    X = np.array(X, dtype=np.result_type(A, B, Cm, X), copy=True)
    for l in range(n):
        for j in range(n):
            for k in range(max(0, n - (l + j)), n):
                X[k, j, l] = 0
    Y = np.einsum("lm,kjm->kjl", A, X)
    Z = np.einsum("jm,kml->kjl", B, Y)
    Y = np.einsum("km,mjl->kjl", Cm, Z)
    return diagtrav(Y)


# ---- structured-kind `+` / `-` (numpy restatement of /root/reference/src/special.jl:75-222) ---------------
# A structured operand is a tuple (kind, dl, d, du) with kind in {"D","BU","BL","S","T"} and None for a diagonal
# the kind does not have; "S" carries ev as both dl and du.  UniformScaling is ("I", lam).
def _sp_join(ka, kb):
    """Return kind of `A +/- B`: Diagonal is absorbed (special.jl:75-142), equal kinds are kept
    (src/tridiag.jl:206-207,:571-572; src/bidiag.jl:330-346 same uplo), anything else is Tridiagonal
    (special.jl:117-178; src/bidiag.jl:334-336 opposite uplo)."""
    if ka == "D":
        return kb
    if kb == "D":
        return ka
    return ka if ka == kb else "T"


def special_axpby(sa, A, sb, B):
    """`sa*A + sb*B` on diagonals for sa, sb in {+1,-1}: every method of special.jl:75-178 is an instance
    (e.g. :92-95 `Diagonal - Bidiagonal` = Bidiagonal(A.diag - B.dv, -B.ev, B.uplo)).  With B = ("I", lam) or
    A = ("I", lam) it is the UniformScaling family :180-222 (`newd = A.d .+ B.lam`, `A.lam .- B.d`, `-B.ev`)."""
    def diag_of(X, k, n):
        v = X[1 + k]
        return None if v is None else np.asarray(v)[: (n if k == 1 else max(n - 1, 0))]

    if A[0] == "I" or B[0] == "I":
        lam, sl, X, sx = (A[1], sa, B, sb) if A[0] == "I" else (B[1], sb, A, sa)
        n = len(X[2])
        out = []
        for k in range(3):
            v = diag_of(X, k, n)
            if v is None:
                out.append(None)
            elif k == 1:
                out.append(sx * v + sl * lam)
            else:
                out.append(sx * v)
        return (X[0], out[0], out[1], out[2])
    n = len(A[2])
    kind = _sp_join(A[0], B[0])
    want = {"D": (0, 1, 0), "BU": (0, 1, 1), "BL": (1, 1, 0), "S": (1, 1, 1), "T": (1, 1, 1)}[kind]
    out = []
    for k in range(3):
        a, b = diag_of(A, k, n), diag_of(B, k, n)
        if not want[k]:
            out.append(None)
        elif a is not None and b is not None:
            out.append(sa * a + sb * b)
        elif a is not None:
            out.append(sa * a)
        elif b is not None:
            out.append(sb * b)
        else:
            out.append(np.zeros(n if k == 1 else max(n - 1, 0), dtype=np.asarray(A[2]).dtype))
    return (kind, out[0], out[1], out[2])


def special_dense(X):
    """`Matrix(A)` of a structured tuple (element definitions src/tridiag.jl:301-314,:471-484; src/bidiag.jl:111-124)."""
    _, dl, d, du = X
    n = len(d)
    M = np.diag(np.asarray(d))
    if n > 1:
        if dl is not None:
            M = M + np.diag(np.asarray(dl)[: n - 1], -1)
        if du is not None:
            M = M + np.diag(np.asarray(du)[: n - 1], 1)
    return M
