"""GPU parity: Tridiagonal / SymTridiagonal / Bidiagonal mul! (mat-vec and mat-mat) vs the oracle."""
import numpy as np
import pytest
import torch

from tests._util import tol_rows

pytestmark = pytest.mark.gpu
NP = {torch.float64: np.float64, torch.float32: np.float32}
SIZES = [1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 1000, 4097, 100003, 1 << 20]


def _mk(oracle, n, dt, seed):
    npdt = NP[dt]
    d = oracle.fill_uniform(seed, n, npdt)
    dl = oracle.fill_uniform(seed + 1, max(n - 1, 0), npdt)
    du = oracle.fill_uniform(seed + 2, max(n - 1, 0), npdt)
    return dl, d, du


def _t(a):
    return torch.from_numpy(a).cuda()


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("kind", ["tri", "sym", "symlong", "biU", "biL"])
def test_structured_matvec(L, oracle, dt, kind):
    npdt = NP[dt]
    for n in SIZES:
        dl, d, du = _mk(oracle, n, dt, 40 + n % 7)
        if kind == "tri":
            A = L.Tridiagonal(_t(dl), _t(d), _t(du))
            o = (dl, d, du)
        elif kind == "sym":
            A = L.SymTridiagonal(_t(d), _t(du))
            o = (du, d, du)
        elif kind == "symlong":  # len(ev) == n is allowed; only ev[:n-1] is used (src/tridiag.jl:11)
            ev = oracle.fill_uniform(3, n, npdt)
            A = L.SymTridiagonal(_t(d), _t(ev))
            o = (ev[: n - 1], d, ev[: n - 1])
        elif kind == "biU":
            A = L.Bidiagonal(_t(d), _t(du), "U")
            o = (None, d, du)
        else:
            A = L.Bidiagonal(_t(d), _t(dl), "L")
            o = (dl, d, None)
        odl, od, odu = o
        if n == 1:
            odl = odu = None
        for alpha, beta in [(1.0, 0.0), (2.0, 1.0)]:
            x = oracle.fill_uniform(50, n, npdt)
            y0 = oracle.fill_uniform(51, n, npdt)
            want = oracle.tridiag_mv(n, odl, od, odu, x, alpha, beta, y0)
            y = _t(y0.copy())
            L.mul_(y, A, _t(x), alpha, beta)
            err = np.max(np.abs(y.cpu().numpy() - want))
            assert err <= tol_rows(npdt, 3, 1.0, 1.0, alpha, beta, 1.0), (kind, n, alpha, beta, err)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_structured_matmat_and_transposes(L, oracle, dt):
    npdt = NP[dt]
    for n in (2, 64, 1000, 4100):
        dl, d, du = _mk(oracle, n, dt, 7)
        X = np.asfortranarray(oracle.fill_uniform(8, n * 3, npdt).reshape(n, 3, order="F"))
        A = L.Tridiagonal(_t(dl), _t(d), _t(du))
        for At, (a, b) in ((A, (dl, du)), (A.T, (du, dl))):          # transpose = swap dl,du
            want = oracle.tridiag_mv(n, a, d, b, X)
            Xg = torch.from_numpy(X).cuda()                            # column-major view (stride (1,n))
            got = (At @ Xg).cpu().numpy()
            assert np.max(np.abs(got - want)) <= tol_rows(npdt, 3, 1.0, 1.0)
            Xr = torch.from_numpy(np.ascontiguousarray(X)).cuda()      # row-major input also accepted
            got = (At @ Xr).cpu().numpy()
            assert np.max(np.abs(got - want)) <= tol_rows(npdt, 3, 1.0, 1.0)
        B = L.Bidiagonal(_t(d), _t(du), "U")
        want = oracle.tridiag_mv(n, du, d, None, X)                    # B' is lower with the same ev
        assert np.max(np.abs((B.T @ torch.from_numpy(X).cuda()).cpu().numpy() - want)) <= tol_rows(npdt, 2, 1.0, 1.0)


def test_structured_kats(L, kats):
    """The reference's own KATs for the structured types, run on the GPU."""
    k = next(k for k in kats if k["kind"] == "symtridiag_mv")           # test/test_tridiag.jl:339-341
    T = L.SymTridiagonal(torch.tensor(k["dv"], dtype=torch.float64).cuda(), torch.tensor(k["ev"], dtype=torch.float64).cuda())
    assert (T @ torch.tensor(k["x"], dtype=torch.float64).cuda()).cpu().tolist() == k["expect"]
    k = next(k for k in kats if k["kind"] == "symtridiag_empty")        # test/test_tridiag.jl:342
    T = L.SymTridiagonal(torch.ones(0, dtype=torch.float64).cuda(), torch.ones(0, dtype=torch.float64).cuda())
    out = T @ torch.ones((0, 2), dtype=torch.float64).cuda()
    assert tuple(out.shape) == (0, 2)
    for kind, mk in (("dense_tridiagonal", lambda k: L.Tridiagonal(*(torch.tensor(k[s], dtype=torch.float64).cuda() for s in ("dl", "d", "du")))),
                     ("dense_symtridiagonal", lambda k: L.SymTridiagonal(*(torch.tensor(k[s], dtype=torch.float64).cuda() for s in ("dv", "ev")))),
                     ("dense_bidiagonal", lambda k: L.Bidiagonal(*(torch.tensor(k[s], dtype=torch.float64).cuda() for s in ("dv", "ev")), k["uplo"]))):
        k = next(k for k in kats if k["kind"] == kind)
        A = mk(k)
        n = len(k["expect"])
        dense = A @ torch.eye(n, dtype=torch.float64).cuda()
        assert dense.cpu().tolist() == k["expect"], k["src"]
        assert A.to_dense().cpu().tolist() == k["expect"]


def test_structured_mul_errors(L):
    """test/test_tridiag.jl:272-279: DimensionMismatch on mis-shaped mul!."""
    n = 6
    z = lambda *s: torch.zeros(*s, dtype=torch.float64, device="cuda")
    A = L.Tridiagonal(z(n - 1), z(n), z(n - 1))
    B = L.SymTridiagonal(z(n), z(n - 1))
    Cnn, Cnm, Cmn = z(n, n), z(n, n + 1), z(n + 1, n)
    for args in [(Cnn, A, Cnm), (Cnn, A, Cmn), (Cnn, B, Cmn), (Cmn, B, Cnn), (Cnm, B, Cnn)]:
        with pytest.raises(L.DimensionMismatch):
            L.mul_(*args)


def test_structured_shuffle_kernel_matvec(L, oracle):
    """The warp-shuffle kernel (mat-mat path) forced onto mat-vec inputs == the TMA-ring kernel."""
    h = L.handle_for(0)
    for n in (1, 2, 65, 4097, 300001):
        dl, d, du = _mk(oracle, n, torch.float64, 3)
        x = oracle.fill_uniform(9, n)
        for (a, b, A) in ((dl, du, L.Tridiagonal(_t(dl), _t(d), _t(du))), (du, du, L.SymTridiagonal(_t(d), _t(du))),
                          (None, du, L.Bidiagonal(_t(d), _t(du), "U")), (dl, None, L.Bidiagonal(_t(d), _t(dl), "L"))):
            want = oracle.tridiag_mv(n, a if n > 1 else None, d, b if n > 1 else None, x)
            ring = (A @ _t(x)).cpu().numpy()
            h.set_tuning(tri_force_shuffle=1)
            try:
                shf = (A @ _t(x)).cpu().numpy()
            finally:
                h.set_tuning(tri_force_shuffle=0)
            assert np.max(np.abs(ring - want)) <= tol_rows(np.float64, 3, 1.0, 1.0)
            assert np.max(np.abs(shf - want)) <= tol_rows(np.float64, 3, 1.0, 1.0)
    for tune in ({"tri_tile": 256, "tri_stages": 1, "tri_ctas_per_sm": 1}, {"tri_tile": 2048, "tri_stages": 3, "tri_ctas_per_sm": 2}):
        h.set_tuning(**tune)
        try:
            n = 700001
            dl, d, du = _mk(oracle, n, torch.float64, 5)
            x = oracle.fill_uniform(9, n)
            got = (L.Tridiagonal(_t(dl), _t(d), _t(du)) @ _t(x)).cpu().numpy()
        finally:
            h.set_tuning(**{k: 0 for k in tune})
        assert np.max(np.abs(got - oracle.tridiag_mv(n, dl, d, du, x))) <= tol_rows(np.float64, 3, 1.0, 1.0)


def test_structured_scalar_fallback_matches(L, oracle):
    n = 50001
    dl, d, du = _mk(oracle, n, torch.float64, 3)
    x = oracle.fill_uniform(9, n)
    want = oracle.tridiag_mv(n, dl, d, du, x)
    A = L.Tridiagonal(_t(dl), _t(d), _t(du))
    h = L.handle_for(0)
    h.set_tuning(tri_force_scalar=1)
    try:
        got = (A @ _t(x)).cpu().numpy()
    finally:
        h.set_tuning(tri_force_scalar=0)
    assert np.max(np.abs(got - want)) <= tol_rows(np.float64, 3, 1.0, 1.0)
    # unaligned views
    buf = _t(oracle.fill_uniform(1, n + 1))
    xv = buf[1:]
    got = (A @ xv).cpu().numpy()
    want = oracle.tridiag_mv(n, dl, d, du, xv.cpu().numpy())
    assert np.max(np.abs(got - want)) <= tol_rows(np.float64, 3, 1.0, 1.0)


def test_tridiag_row_ranges(L, oracle):
    """lbm_?tridiag_mv_rows: only rows [a,b) are produced, written at y[i-a] (used by row shards)."""
    import ctypes as C
    h = L.handle_for(0)
    n = 70003
    dl, d, du = _mk(oracle, n, torch.float64, 3)
    x = oracle.fill_uniform(9, n)
    want = oracle.tridiag_mv(n, dl, d, du, x, 2.0, 0.0)
    g = [_t(a) for a in (dl, d, du, x)]
    for (a, b) in ((0, n), (1, n - 1), (1, 1025), (1023, 1024), (4096, 4096), (5000, n), (33333, 66001)):
        y = torch.full((max(b - a, 1),), float("nan"), dtype=torch.float64, device="cuda")
        rc = L.lib().lbm_dtridiag_mv_rows(h.ptr, n, C.c_void_p(g[0].data_ptr()), C.c_void_p(g[1].data_ptr()),
                                          C.c_void_p(g[2].data_ptr()), 2.0, C.c_void_p(g[3].data_ptr()), 0.0,
                                          C.c_void_p(y.data_ptr()), a, b)
        assert rc == 0
        torch.cuda.synchronize()
        if b > a:
            assert np.max(np.abs(y.cpu().numpy()[: b - a] - want[a:b])) <= tol_rows(np.float64, 3, 1.0, 1.0, 2.0)
    assert L.lib().lbm_dtridiag_mv_rows(h.ptr, n, None, C.c_void_p(g[1].data_ptr()), None, 1.0,
                                        C.c_void_p(g[3].data_ptr()), 0.0, C.c_void_p(g[3].data_ptr()), 5, 3) == -13


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_dot_and_sum_reductions(L, oracle, dt):
    """dot(x, A, y) (src/tridiag.jl:665-682) and Base._sum(A[, dims]) (src/tridiag.jl:594-660, src/bidiag.jl:396-435)
    as single-pass device reductions, all three kinds.  dims sums add <= 3 terms: tolerance 4*eps*3; scalar
    reductions are compared with a float64 / math.fsum evaluation under 4*eps*(sum of |terms|) (tree summation is
    at least as accurate as the reference's sequential loop, whose own bound is n*eps*(sum of |terms|))."""
    import math
    npdt = np.float64 if dt == torch.float64 else np.float32
    eps = float(np.finfo(npdt).eps)
    for n in (0, 1, 2, 3, 257, 100003, 1 << 20):
        d = oracle.fill_uniform(21, n, npdt) - 0.5
        lo = oracle.fill_uniform(22, max(n - 1, 0), npdt) - 0.5
        up = oracle.fill_uniform(23, max(n - 1, 0), npdt) - 0.5
        x = oracle.fill_uniform(24, n, npdt) - 0.5
        y = oracle.fill_uniform(25, n, npdt) - 0.5
        g = lambda a: torch.from_numpy(a).cuda()
        kinds = [
            (L.Tridiagonal(g(lo), g(d), g(up)), lo, up),
            (L.SymTridiagonal(g(d), g(up)), up, up),
            (L.Bidiagonal(g(d), g(up), "U"), None, up),
            (L.Bidiagonal(g(d), g(lo), "L"), lo, None),
        ]
        for A, l_np, u_np in kinds:
            # dims sums
            for dims in (1, 2):
                got = A.sum(dims).cpu().numpy()
                want = oracle.tridiag_sum(n, l_np, d, u_np, dims)
                assert got.shape == want.shape
                assert np.max(np.abs(got - want), initial=0.0) <= 4 * eps * 3
            # scalar sum
            terms = [d.astype(np.float64)] + [t.astype(np.float64) for t in (l_np, u_np) if t is not None]
            want = math.fsum(np.concatenate(terms)) if n else 0.0
            bound = 4 * eps * float(sum(np.abs(t).sum() for t in terms)) + 1e-300
            assert abs(float(A.sum()) - want) <= bound, (n, type(A).__name__)
            # dot(x, A, y)
            zl = np.zeros(max(n - 1, 0), npdt)
            l64 = (zl if l_np is None else l_np).astype(np.float64)
            u64 = (zl if u_np is None else u_np).astype(np.float64)
            want = oracle.tridiag_dot(x.astype(np.float64), l64, d.astype(np.float64), u64, y.astype(np.float64)) if n else 0.0
            x64, y64 = np.abs(x.astype(np.float64)), np.abs(y.astype(np.float64))
            mag = float((np.abs(d) * x64 * y64).sum()) + (float((np.abs(l64) * x64[1:] * y64[:-1]).sum() + (np.abs(u64) * x64[:-1] * y64[1:]).sum()) if n > 1 else 0.0)
            assert abs(float(A.dot(g(x), g(y))) - want) <= 4 * eps * 3 * mag + 1e-300, (n, type(A).__name__)
    with pytest.raises(L.DimensionMismatch):
        L.SymTridiagonal(torch.ones(4, device="cuda"), torch.ones(3, device="cuda")).dot(torch.ones(5, device="cuda"), torch.ones(4, device="cuda"))
