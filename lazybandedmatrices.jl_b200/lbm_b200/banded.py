"""`BandedMatrix` / `brand` -- the GPU-resident storage the kernels consume.

Host-side mirror of BandedMatrices.jl's type ([up], un-vendored; imported by the reference at
/root/reference/src/LazyBandedMatrices.jl:21): `data` is the `(l+u+1) x n` column-major
array with `A[i,j] = data[u+i-j, j]`.  A torch CUDA tensor of shape `(n, l+u+1)` (C-contiguous)
has exactly that memory layout, so `data.data_ptr()` is what the Julia shim would pass.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .errors import DimensionMismatch, LbmArgumentError

_PFX = {torch.float64: "d", torch.float32: "s"}


def _pfx(dtype):
    try:
        return _PFX[dtype]
    except KeyError:
        raise LbmArgumentError(f"unsupported element type {dtype}: the B200 path covers Float64/Float32")


def _ft(dtype):
    return _lib.f64 if dtype == torch.float64 else _lib.f32


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def fill_uniform(out: torch.Tensor, seed: int, offset: int = 0) -> torch.Tensor:
    """Counter-based U[0,1) stream, generated on the device."""
    h = _lib.handle_of(out)
    fn = getattr(_lib.lib(), f"lbm_{_pfx(out.dtype)}fill_uniform")
    assert out.is_contiguous()
    h.check(fn(h.ptr, seed & (2**64 - 1), offset, out.numel(), _ptr(out)), "fill_uniform")
    return out


class BandedMatrix:
    """m x n banded matrix with bandwidths (l, u) in diagonal-major storage."""

    def __init__(self, data: torch.Tensor, m: int, l: int, u: int):
        if data.dim() != 2 or data.shape[1] < l + u + 1:
            raise LbmArgumentError("data must be (n, lda) with lda >= l+u+1 (column-major (l+u+1) x n)")
        if not data.is_contiguous():
            data = data.contiguous()
        self.data = data
        self.m = int(m)
        self.n = int(data.shape[0])
        self.l = int(l)
        self.u = int(u)

    # ---- reference-API surface -------------------------------------------------------
    @property
    def lda(self):
        return int(self.data.shape[1])

    @property
    def shape(self):
        return (self.m, self.n)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def device(self):
        return self.data.device

    def size(self, d=None):
        return self.shape if d is None else self.shape[d - 1]

    def bandwidths(self):
        return (self.l, self.u)

    @property
    def T(self):
        return TransposedBanded(self)

    def to_dense(self) -> torch.Tensor:
        """Dense copy (tests / small cases only; masks out-of-matrix slots by index)."""
        A = torch.zeros(self.m, self.n, dtype=self.dtype, device=self.device)
        for r in range(self.l + self.u + 1):
            k = self.u - r  # diagonal offset j - i
            lo, hi = max(0, k), min(self.n, self.m + k)
            j = torch.arange(lo, max(lo, hi), device=self.device)
            if j.numel():
                A[j - k, j] = self.data[j, r]
        return A

    def __matmul__(self, other):
        from .linalg import matmul
        return matmul(self, other)

    def __repr__(self):
        return f"BandedMatrix({self.m}x{self.n}, bandwidths=({self.l},{self.u}), {self.dtype}, {self.device})"


class TransposedBanded:
    """`A'` of a BandedMatrix: applies through gbmv with trans='T' (no data movement)."""

    def __init__(self, parent: BandedMatrix):
        self.parent = parent

    @property
    def shape(self):
        return (self.parent.n, self.parent.m)

    @property
    def dtype(self):
        return self.parent.dtype

    @property
    def device(self):
        return self.parent.device

    def bandwidths(self):
        return (self.parent.u, self.parent.l)

    @property
    def T(self):
        return self.parent

    def to_dense(self):
        return self.parent.to_dense().t()

    def __matmul__(self, other):
        from .linalg import matmul
        return matmul(self, other)


def brand(m: int, n: int, l: int, u: int, dtype=torch.float64, device="cuda", seed: int = 1) -> BandedMatrix:
    """`brand(m,n,l,u)`: every one of the (l+u+1)*n storage slots is U[0,1), including the
    out-of-matrix corners ([up] BandedMatrices semantics)."""
    data = torch.empty((n, l + u + 1), dtype=dtype, device=device)
    if data.numel():
        fill_uniform(data, seed)
    return BandedMatrix(data, m, l, u)


def banded_from_diagonals(n: int, diags: dict, dtype=torch.float64, device="cuda") -> BandedMatrix:
    """`BandedMatrix(k1 => v1, ...)` ([up]; used at /root/reference/test/test_blockkron.jl:96,299)."""
    ks = list(diags)
    l = max(0, -min(ks))
    u = max(0, max(ks))
    data = torch.zeros((n, l + u + 1), dtype=dtype, device=device)
    for k, v in diags.items():
        v = torch.as_tensor(v, dtype=dtype, device=device).expand(n - abs(k))
        if k >= 0:
            data[k:, u - k] = v
        else:
            data[: n + k, u - k] = v
    return BandedMatrix(data, n, l, u)


# ---- kernels ---------------------------------------------------------------------------------
def gbmv_(y, A: BandedMatrix, x, alpha=1.0, beta=0.0, trans="N"):
    """`mul!(y, A, x, alpha, beta)` for a BandedMatrix -> lbm_?gbmv (replaces BandedMatrices.gbmv!)."""
    m, n = A.shape
    nin, nout = (m, n) if trans != "N" else (n, m)
    if x.shape[0] != nin:
        raise DimensionMismatch(
            f"second entry of size(A)={(nout, nin)} and first entry of size(B) = {tuple(x.shape)} must match.")
    if y.shape[0] != nout:
        raise DimensionMismatch(f"sizes size(A)={(nout, nin)} and size(C) = {tuple(y.shape)} must match at first entry.")
    if x.dim() != y.dim() or (x.dim() == 2 and x.shape[1] != y.shape[1]):
        raise DimensionMismatch(f"sizes size(B)={tuple(x.shape)} and size(C) = {tuple(y.shape)} must match at first second entry.")
    h = _lib.handle_of(A.data)
    fn = getattr(_lib.lib(), f"lbm_{_pfx(A.dtype)}gbmv")
    ft = _ft(A.dtype)
    if x.dim() == 1:
        xcc = x if x.is_contiguous() else x.contiguous()
        ycc = y if y.is_contiguous() else y.contiguous()
        h.check(fn(h.ptr, trans.encode(), m, n, A.l, A.u, ft(alpha), _ptr(A.data), A.lda, _ptr(xcc), ft(beta),
                   _ptr(ycc)), "lbm_gbmv")
        if ycc is not y:
            y.copy_(ycc)
        return y
    # mat-mat: X, Y are torch (rows, nrhs) tensors; the ABI wants column-major storage with a leading dimension,
    # i.e. the transposed tensor with unit inner stride (a padded outer stride is passed through as ldx / ldy)
    nrhs = x.shape[1]

    def colmajor(t, rows):
        tt = t.t()
        ok = tt.stride(1) == 1 and (nrhs == 1 or tt.stride(0) >= max(1, rows))
        return (tt, True) if ok else (tt.contiguous(), False)

    xt, _ = colmajor(x, nin)
    yt, y_direct = colmajor(y, nout)
    ldx = xt.stride(0) if nrhs > 1 else max(1, nin)
    ldy = yt.stride(0) if nrhs > 1 else max(1, nout)
    fm = getattr(_lib.lib(), f"lbm_{_pfx(A.dtype)}gbmv_multi")
    h.check(fm(h.ptr, trans.encode(), m, n, A.l, A.u, ft(alpha), _ptr(A.data), A.lda, _ptr(xt), ldx, ft(beta),
               _ptr(yt), ldy, nrhs), "lbm_gbmv_multi")
    if not y_direct:
        y.copy_(yt.t())
    return y


def gbmm(A: BandedMatrix, B: BandedMatrix, alpha=1.0, out: BandedMatrix = None, beta=0.0) -> BandedMatrix:
    """Materialise `A*B` -> BandedMatrix with band (l1+l2, u1+u2) (lbm_?gbmm)."""
    if A.n != B.m:
        raise DimensionMismatch(
            f"second entry of size(A)={A.shape} and first entry of size(B) = {B.shape} must match.")
    m, k, n = A.m, A.n, B.n
    lc, uc = A.l + B.l, A.u + B.u
    if out is None:
        out = BandedMatrix(torch.empty((n, lc + uc + 1), dtype=A.dtype, device=A.device), m, lc, uc)
        beta = 0.0
    elif out.shape != (m, n):
        raise DimensionMismatch(f"sizes size(A)={A.shape} and size(C) = {out.shape} must match at first entry.")
    h = _lib.handle_of(A.data)
    fn = getattr(_lib.lib(), f"lbm_{_pfx(A.dtype)}gbmm")
    ft = _ft(A.dtype)
    h.check(fn(h.ptr, m, n, k, ft(alpha), A.l, A.u, _ptr(A.data), A.lda, B.l, B.u, _ptr(B.data), B.lda, ft(beta),
               out.l, out.u, _ptr(out.data), out.lda), "lbm_gbmm")
    return out
