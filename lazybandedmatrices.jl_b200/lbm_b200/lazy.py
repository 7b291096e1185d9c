"""`ApplyMatrix(*, A, B, ...)`: lazy banded products, their `bandwidths`, `materialize` and `mul!`.

Host-side mirror of the LazyArrays surface the reference imports
(/root/reference/src/LazyBandedMatrices.jl:18-20; shown in README.md:8-26):

  * `bandwidths(M) = min.(size(M) .- 1, (sum l_i, sum u_i))`                       
  * `BandedMatrix(M)` / `materialize(M)`: pairwise left-to-right banded products   (item 3)
  * `mul!(y, M, x)`: flattened right-to-left `A*(B*(...x))` with k-1 temporaries  (item 4; the
    same rule the reference relies on at src/blockkron.jl:208-214, test/test_blockkron.jl:345-348)
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .banded import BandedMatrix, _pfx, _ptr, gbmm
from .errors import DimensionMismatch, LbmArgumentError


class ApplyMatrix:
    def __init__(self, f, *args):
        if f not in ("*", "mul"):
            raise LbmArgumentError("only ApplyMatrix(*, ...) is on the B200 path")
        if len(args) < 1:
            raise LbmArgumentError("ApplyMatrix(*) needs factors")
        for a, b in zip(args[:-1], args[1:]):
            if a.shape[1] != b.shape[0]:
                raise DimensionMismatch(
                    f"second entry of size(A)={tuple(a.shape)} and first entry of size(B) = {tuple(b.shape)} must match.")
        self.f = "*"
        self.args = tuple(args)

    @property
    def shape(self):
        return (self.args[0].shape[0], self.args[-1].shape[1])

    @property
    def dtype(self):
        return self.args[0].dtype

    @property
    def device(self):
        return self.args[0].device

    def bandwidths(self):
        m, n = self.shape
        kl = (C.c_int64 * len(self.args))(*[a.bandwidths()[0] for a in self.args])
        ku = (C.c_int64 * len(self.args))(*[a.bandwidths()[1] for a in self.args])
        ol, ou = C.c_int64(), C.c_int64()
        rc = _lib.lib().lbm_prod_bandwidths(m, n, len(self.args), kl, ku, C.byref(ol), C.byref(ou))
        if rc:
            raise LbmArgumentError(f"lbm_prod_bandwidths: argument {-rc} invalid")
        return (int(ol.value), int(ou.value))

    def materialize(self) -> BandedMatrix:
        """`BandedMatrix(ApplyMatrix(*, ...))`: pairwise, left to right."""
        out = self.args[0]
        for nxt in self.args[1:]:
            out = gbmm(_as_banded(out), _as_banded(nxt))
        return out

    def mul_(self, y, x, alpha=1.0, beta=0.0):
        from .linalg import mul_ as _mul
        if x.shape[0] != self.shape[1]:
            raise DimensionMismatch(
                f"second entry of size(A)={self.shape} and first entry of size(B) = {tuple(x.shape)} must match.")
        if y.shape[0] != self.shape[0]:
            raise DimensionMismatch(f"sizes size(A)={self.shape} and size(C) = {tuple(y.shape)} must match at first entry.")
        # all-banded product applied to a vector: one C call (lbm_?gbmv_chain), temporaries in the handle's scratch
        if (x.dim() == 1 and len(self.args) > 1 and all(isinstance(a, BandedMatrix) for a in self.args)
                and all(a.dtype == x.dtype for a in self.args)):
            nf = len(self.args)
            key = tuple(a.data.data_ptr() for a in self.args)
            desc = getattr(self, "_chain_desc", None)
            if desc is None or desc[0] != key:      # factor description in C arrays, built once per set of buffers
                arr = lambda vals: (C.c_int64 * nf)(*vals)  # noqa: E731
                desc = (key, arr([a.m for a in self.args]), arr([a.n for a in self.args]), arr([a.l for a in self.args]),
                        arr([a.u for a in self.args]), (C.c_void_p * nf)(*key), arr([a.lda for a in self.args]))
                self._chain_desc = desc
            h = _lib.handle_of(x)
            fn = getattr(_lib.lib(), f"lbm_{_pfx(x.dtype)}gbmv_chain")
            ft = _lib.f64 if x.dtype == torch.float64 else _lib.f32
            xc = x if x.is_contiguous() else x.contiguous()
            yc = y if y.is_contiguous() else y.contiguous()
            h.check(fn(h.ptr, nf, desc[1], desc[2], desc[3], desc[4], desc[5], desc[6], ft(alpha), _ptr(xc), ft(beta), _ptr(yc)),
                    "lbm_gbmv_chain")
            if yc is not y:
                y.copy_(yc)
            return y
        v = x
        for a in reversed(self.args[1:]):
            t = torch.empty((a.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            _mul(t, a, v)
            v = t
        return _mul(y, self.args[0], v, alpha, beta)

    def __matmul__(self, x):
        y = torch.empty((self.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        return self.mul_(y, x)


class BroadcastMatrix:
    """`BroadcastArray(+, A, B, ...)` / `BroadcastArray(-, A, B)` of banded operands (the
    `BroadcastBandedLayout` the reference imports at /root/reference/src/LazyBandedMatrices.jl:20;
    used for the Laplacian sum at test/test_blockkron.jl:303,345).  `bandwidths` is the union band;
    `mul!` applies term by term, accumulating into y with beta = 1 (`A*x + B*x`), which is how the
    reference's `Mul{BroadcastBandedLayout{+}}` simplifies."""

    def __init__(self, f, *args):
        if f not in ("+", "-"):
            raise LbmArgumentError("only BroadcastArray(+/-, ...) of banded operands is on the B200 path")
        if f == "-" and len(args) != 2:
            raise LbmArgumentError("BroadcastArray(-) takes two operands")
        for a in args[1:]:
            if tuple(a.shape) != tuple(args[0].shape):
                raise DimensionMismatch(f"arrays could not be broadcast to a common size: {tuple(args[0].shape)} and {tuple(a.shape)}")
        self.f = f
        self.args = tuple(args)

    @property
    def shape(self):
        return tuple(self.args[0].shape)

    def bandwidths(self):
        ls, us = zip(*[a.bandwidths() for a in self.args])
        return (max(ls), max(us))

    def mul_(self, y, x, alpha=1.0, beta=0.0):
        from .linalg import mul_ as _mul
        sign = [1.0] + [(-1.0 if self.f == "-" else 1.0)] * (len(self.args) - 1)
        if x.shape[0] != self.shape[1]:
            raise DimensionMismatch(
                f"second entry of size(A)={self.shape} and first entry of size(B) = {tuple(x.shape)} must match.")
        if y.shape[0] != self.shape[0]:
            raise DimensionMismatch(f"sizes size(A)={self.shape} and size(C) = {tuple(y.shape)} must match at first entry.")
        if (len(self.args) == 2 and all(isinstance(a, BandedMatrix) for a in self.args) and x.dim() == 1
                and x.is_contiguous() and y.is_contiguous() and self.args[0].dtype == self.args[1].dtype == x.dtype):
            # fused: one pass over x and y, the sum is never formed (lbm_?gbmv2)
            A1, A2 = self.args
            h = _lib.handle_of(A1.data)
            ft = _lib.f64 if A1.dtype == torch.float64 else _lib.f32
            fn = getattr(_lib.lib(), f"lbm_{_pfx(A1.dtype)}gbmv2")
            h.check(fn(h.ptr, A1.m, A1.n, ft(alpha * sign[0]), A1.l, A1.u, _ptr(A1.data), A1.lda, ft(alpha * sign[1]),
                       A2.l, A2.u, _ptr(A2.data), A2.lda, _ptr(x), ft(beta), _ptr(y)), "lbm_gbmv2")
            return y
        for k, (a, sg) in enumerate(zip(self.args, sign)):
            _mul(y, a, x, alpha * sg, beta if k == 0 else 1.0)
        return y

    def __matmul__(self, x):
        y = torch.empty((self.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        return self.mul_(y, x)


def _as_banded(a) -> BandedMatrix:
    if isinstance(a, BandedMatrix):
        return a
    if hasattr(a, "as_banded"):
        return a.as_banded()
    raise LbmArgumentError(f"cannot materialise a factor of type {type(a).__name__} as BandedMatrix")


def chain3_batched(A: torch.Tensor, B: torch.Tensor, Cm: torch.Tensor, kl, ku) -> torch.Tensor:
    """Batched `A_b*B_b*C_b` materialisation, fused (the intermediate band stays on chip).
    Factor tensors are (batch, n, kl_i+ku_i+1): per item the column-major band array."""
    batch, n = A.shape[0], A.shape[1]
    for t, l, u in zip((A, B, Cm), kl, ku):
        if t.shape[0] != batch or t.shape[1] != n or t.shape[2] != l + u + 1 or not t.is_contiguous():
            raise DimensionMismatch("factor tensors must be contiguous (batch, n, kl+ku+1)")
    ldd = sum(kl) + sum(ku) + 1
    D = torch.empty((batch, n, ldd), dtype=A.dtype, device=A.device)
    h = _lib.handle_of(A)
    fn = getattr(_lib.lib(), f"lbm_{_pfx(A.dtype)}gbmm_chain3_batched")
    klc = (C.c_int64 * 3)(*kl)
    kuc = (C.c_int64 * 3)(*ku)
    done = 0
    while done < batch:  # the kernel's grid.y covers <= 65535 items per call
        nb = min(batch - done, 65535)
        h.check(fn(h.ptr, nb, n, klc, kuc, _ptr(A[done]), A[0].numel(), _ptr(B[done]), B[0].numel(),
                   _ptr(Cm[done]), Cm[0].numel(), _ptr(D[done]), D[0].numel()), "lbm_gbmm_chain3_batched")
        done += nb
    return D
