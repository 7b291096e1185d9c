"""`Tridiagonal`, `SymTridiagonal`, `Bidiagonal` with GPU-resident diagonals.

Host-side mirrors of the reference's types (same field names, same constructor checks, same
`bandwidths`), whose `mul!` goes to lbm_?tridiag_mv / lbm_?bidiag_mv:

  Tridiagonal(dl,d,du)    /root/reference/src/tridiag.jl:327-341 (ctor checks :331-340)
  SymTridiagonal(dv,ev)   src/tridiag.jl:6-16 (len(ev) in {n-1, n}, :11)
  Bidiagonal(dv,ev,uplo)  src/bidiag.jl:4-15 (uplo in 'U','L'; :16-18)
"""
from __future__ import annotations

import torch

from . import _lib
from .banded import _ft, _pfx, _ptr
from .errors import DimensionMismatch, LbmArgumentError


def _colmajor(X: torch.Tensor):
    """(tensor with column-major memory, ld).  1-D -> itself."""
    if X.dim() == 1:
        return (X if X.is_contiguous() else X.contiguous()), max(int(X.shape[0]), 1)
    n, nrhs = X.shape
    if X.stride(0) == 1 and (nrhs <= 1 or X.stride(1) >= max(n, 1)):
        return X, int(X.stride(1)) if nrhs > 1 else max(n, 1)
    Xc = X.t().contiguous().t()
    return Xc, max(n, 1)


def _check_mul_sizes(A_shape, x, y):
    """check_A_mul_B!_sizes (/root/reference/src/bidiag.jl:363-377), same messages."""
    nA, mA = A_shape
    nB = x.shape[0]
    mB = x.shape[1] if x.dim() == 2 else 1
    nC = y.shape[0]
    mC = y.shape[1] if y.dim() == 2 else 1
    if nA != nC:
        raise DimensionMismatch(f"sizes size(A)={A_shape} and size(C) = {tuple(y.shape)} must match at first entry.")
    if mA != nB:
        raise DimensionMismatch(
            f"second entry of size(A)={A_shape} and first entry of size(B) = {tuple(x.shape)} must match.")
    if mB != mC or x.dim() != y.dim():
        raise DimensionMismatch(
            f"sizes size(B)={tuple(x.shape)} and size(C) = {tuple(y.shape)} must match at first second entry.")


def _tridiag_mv_(y, n, dl, d, du, x, alpha, beta):
    _check_mul_sizes((n, n), x, y)
    if n == 0 or x.numel() == 0:
        return y
    h = _lib.handle_of(d)
    fn = getattr(_lib.lib(), f"lbm_{_pfx(d.dtype)}tridiag_mv")
    ft = _ft(d.dtype)
    Xc, ldx = _colmajor(x)
    Yc, ldy = _colmajor(y)
    nrhs = x.shape[1] if x.dim() == 2 else 1
    h.check(fn(h.ptr, n, _ptr(dl), _ptr(d), _ptr(du), ft(alpha), _ptr(Xc), ldx, nrhs, ft(beta), _ptr(Yc), ldy),
            "lbm_tridiag_mv")
    if Yc is not y:
        y.copy_(Yc)
    return y


class _Structured:
    def size(self, d=None):
        return self.shape if d is None else self.shape[d - 1]

    def _diagonals(self):
        """(dl, d, du) as the C ABI wants them: None for a diagonal the kind does not have."""
        raise NotImplementedError

    def dot(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """`dot(x, A, y)` = x' A y as a 0-d device tensor (/root/reference/src/tridiag.jl:665-682); one pass."""
        if not (x.shape[0] == self.n == y.shape[0]) or x.dim() != 1 or y.dim() != 1:
            raise DimensionMismatch(f"dot(x, A, y): lengths {tuple(x.shape)}, {self.shape}, {tuple(y.shape)}")
        dl, d, du = self._diagonals()
        out = torch.empty((), dtype=d.dtype, device=d.device)
        h = _lib.handle_of(d)
        fn = getattr(_lib.lib(), f"lbm_{_pfx(d.dtype)}tridiag_dot")
        h.check(fn(h.ptr, self.n, _ptr(dl), _ptr(d), _ptr(du), _ptr(x.contiguous()), _ptr(y.contiguous()), _ptr(out)),
                "lbm_tridiag_dot")
        return out

    def sum(self, dims=None) -> torch.Tensor:
        """`sum(A)` (0-d), `sum(A; dims=1)` (1 x n) or `sum(A; dims=2)` (n x 1): Base._sum of the reference
        (src/tridiag.jl:594-660, src/bidiag.jl:396-435)."""
        if dims not in (None, 1, 2):
            raise LbmArgumentError("dims must be None, 1 or 2 on the B200 path")
        dl, d, du = self._diagonals()
        out = torch.empty(() if dims is None else (self.n,), dtype=d.dtype, device=d.device)
        h = _lib.handle_of(d)
        fn = getattr(_lib.lib(), f"lbm_{_pfx(d.dtype)}tridiag_sum")
        h.check(fn(h.ptr, self.n, _ptr(dl), _ptr(d), _ptr(du), 0 if dims is None else dims, _ptr(out)), "lbm_tridiag_sum")
        if dims is None:
            return out
        return out.reshape(1, self.n) if dims == 1 else out.reshape(self.n, 1)

    def solve(self, b: torch.Tensor, out: torch.Tensor = None, check: bool = True) -> torch.Tensor:
        """`A \\ b` (vector or n x nrhs matrix): lbm_?tridiag_solve -- the recursive partition method, no pivoting (exact
        for diagonally dominant / SPD operators).  In the reference `\\` reaches LinearAlgebra's tridiagonal LU through
        ArrayLayouts (exercised at test/test_tridiag.jl:345-363).  `check`: raise on a zero pivot (costs one device sync)."""
        if b.shape[0] != self.n or b.dim() not in (1, 2):
            raise DimensionMismatch(f"A has size {self.shape}, b has size {tuple(b.shape)}")
        dl, d, du = self._diagonals()
        if b.dtype != d.dtype:
            raise LbmArgumentError(f"element types differ: {d.dtype} vs {b.dtype}")
        if out is None:
            out = torch.empty_like(b)
        if self.n == 0 or b.numel() == 0:
            return out
        Bc, ldb = _colmajor(b)
        Xc, ldx = _colmajor(out)
        nrhs = b.shape[1] if b.dim() == 2 else 1
        info = torch.zeros(1, dtype=torch.int32, device=d.device)
        h = _lib.handle_of(d)
        fn = getattr(_lib.lib(), f"lbm_{_pfx(d.dtype)}tridiag_solve")
        h.check(fn(h.ptr, self.n, _ptr(dl), _ptr(d), _ptr(du), _ptr(Bc), ldb, nrhs, _ptr(Xc), ldx, _ptr(info)), "lbm_tridiag_solve")
        if Xc is not out:
            out.copy_(Xc)
        if check and int(info.item()) != 0:
            from .errors import SingularException
            raise SingularException("zero pivot in the (unpivoted) tridiagonal elimination")
        return out

    @property
    def shape(self):
        return (self.n, self.n)

    @property
    def dtype(self):
        return self.d.dtype

    @property
    def device(self):
        return self.d.device

    def __matmul__(self, other):
        from .linalg import matmul
        return matmul(self, other)


class Tridiagonal(_Structured):
    def __init__(self, dl: torch.Tensor, d: torch.Tensor, du: torch.Tensor):
        n = d.shape[0]
        # constructor checks of src/tridiag.jl:331-340
        if not (dl.shape[0] == du.shape[0] and (dl.shape[0] == n - 1 or (n == 0 and dl.shape[0] == 0))):
            raise LbmArgumentError(
                f"subdiagonal has wrong length. Has length {dl.shape[0]}, but should be {max(n - 1, 0)}."
                if dl.shape[0] != max(n - 1, 0) else
                f"superdiagonal has wrong length. Has length {du.shape[0]}, but should be {max(n - 1, 0)}.")
        self.dl, self.d, self.du = dl.contiguous(), d.contiguous(), du.contiguous()
        self.n = int(n)

    def bandwidths(self):
        return (1, 1)  # src/tridiag.jl:688

    # accessors of src/tridiag.jl:705-712
    def subdiagonaldata(self):
        return self.dl

    def diagonaldata(self):
        return self.d

    def supdiagonaldata(self):
        return self.du

    @property
    def T(self):
        return Tridiagonal(self.du, self.d, self.dl)  # transpose swaps dl,du (src/tridiag.jl:449-450)

    def to_dense(self):
        A = torch.diag(self.d)
        if self.n > 1:
            A = A + torch.diag(self.dl, -1) + torch.diag(self.du, 1)
        return A

    def _diagonals(self):
        return (self.dl if self.n > 1 else None, self.d, self.du if self.n > 1 else None)

    def mul_(self, y, x, alpha=1.0, beta=0.0):
        return _tridiag_mv_(y, self.n, self.dl if self.n > 1 else None, self.d,
                            self.du if self.n > 1 else None, x, alpha, beta)


class SymTridiagonal(_Structured):
    def __init__(self, dv: torch.Tensor, ev: torch.Tensor):
        n = dv.shape[0]
        if not (ev.shape[0] == n - 1 or ev.shape[0] == n):  # src/tridiag.jl:11
            raise DimensionMismatch(
                f"subdiagonal has wrong length. Has length {ev.shape[0]}, but should be either {n - 1} or {n}.")
        self.dv, self.ev = dv.contiguous(), ev.contiguous()
        self.n = int(n)

    @property
    def d(self):
        return self.dv

    def bandwidths(self):
        return (1, 1)  # src/tridiag.jl:687

    def diagonaldata(self):
        return self.dv

    def supdiagonaldata(self):
        return self.ev  # src/tridiag.jl:708-709

    def subdiagonaldata(self):
        return self.ev

    @property
    def T(self):
        return self

    def to_dense(self):
        A = torch.diag(self.dv)
        if self.n > 1:
            e = self.ev[: self.n - 1]
            A = A + torch.diag(e, -1) + torch.diag(e, 1)
        return A

    def _diagonals(self):
        ev = self.ev if self.n > 1 else None
        return (ev, self.dv, ev)

    def mul_(self, y, x, alpha=1.0, beta=0.0):
        ev = self.ev if self.n > 1 else None
        return _tridiag_mv_(y, self.n, ev, self.dv, ev, x, alpha, beta)  # dl == du: ev read once


class Bidiagonal(_Structured):
    def __init__(self, dv: torch.Tensor, ev: torch.Tensor, uplo):
        uplo = {"U": "U", "L": "L", ":U": "U", ":L": "L"}.get(str(uplo))
        if uplo is None:
            raise LbmArgumentError("uplo argument must be 'U' (upper) or 'L' (lower)")  # src/bidiag.jl:16-18
        n = dv.shape[0]
        if ev.shape[0] != max(n - 1, 0):  # src/bidiag.jl:10
            raise DimensionMismatch(f"length of diagonal vector is {n}, length of off-diagonal is {ev.shape[0]}")
        self.dv, self.ev, self.uplo = dv.contiguous(), ev.contiguous(), uplo
        self.n = int(n)

    @property
    def d(self):
        return self.dv

    def bandwidths(self):
        return (0, 1) if self.uplo == "U" else (1, 0)  # src/bidiag.jl:439

    def diagonaldata(self):
        return self.dv

    def supdiagonaldata(self):
        if self.uplo != "U":
            raise LbmArgumentError("supdiagonaldata of a lower Bidiagonal")
        return self.ev

    def subdiagonaldata(self):
        if self.uplo != "L":
            raise LbmArgumentError("subdiagonaldata of an upper Bidiagonal")
        return self.ev

    @property
    def T(self):
        return Bidiagonal(self.dv, self.ev, "L" if self.uplo == "U" else "U")  # src/bidiag.jl:222-233

    def to_dense(self):
        A = torch.diag(self.dv)
        if self.n > 1:
            A = A + torch.diag(self.ev, 1 if self.uplo == "U" else -1)
        return A

    def _diagonals(self):
        ev = self.ev if self.n > 1 else None
        return (None, self.dv, ev) if self.uplo == "U" else (ev, self.dv, None)

    def mul_(self, y, x, alpha=1.0, beta=0.0):
        _check_mul_sizes((self.n, self.n), x, y)
        if self.n == 0 or x.numel() == 0:
            return y
        h = _lib.handle_of(self.dv)
        fn = getattr(_lib.lib(), f"lbm_{_pfx(self.dtype)}bidiag_mv")
        ft = _ft(self.dtype)
        Xc, ldx = _colmajor(x)
        Yc, ldy = _colmajor(y)
        nrhs = x.shape[1] if x.dim() == 2 else 1
        h.check(fn(h.ptr, self.uplo.encode(), self.n, _ptr(self.dv), _ptr(self.ev if self.n > 1 else None),
                   ft(alpha), _ptr(Xc), ldx, nrhs, ft(beta), _ptr(Yc), ldy), "lbm_bidiag_mv")
        if Yc is not y:
            y.copy_(Yc)
        return y
