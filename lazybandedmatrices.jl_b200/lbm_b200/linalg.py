"""`mul!(y, A, x, alpha, beta)` / `A*x` dispatch over the operand kinds of the hot path.

In the reference this dispatch is ArrayLayouts' `materialize!(::MulAdd{LayA,LayB,LayC})` keyed
on `MemoryLayout` (trait declarations at /root/reference/src/tridiag.jl:685-686,
src/bidiag.jl:438); here it is a type switch onto the C-ABI entry points.
"""
from __future__ import annotations

import torch

from .banded import BandedMatrix, TransposedBanded, gbmm, gbmv_
from .errors import LbmArgumentError


def mul_(y, A, x, alpha=1.0, beta=0.0):
    """In-place `y <- alpha*A*x + beta*y`."""
    if isinstance(A, BandedMatrix):
        return gbmv_(y, A, x, alpha, beta, "N")
    if isinstance(A, TransposedBanded):
        return gbmv_(y, A.parent, x, alpha, beta, "T")
    if hasattr(A, "mul_"):
        return A.mul_(y, x, alpha, beta)
    raise LbmArgumentError(f"mul! is not defined for {type(A).__name__} on the B200 path")


def matmul(A, other):
    """`A*x`, `A*X` (dense right-hand side) or `A*B` (banded x banded -> BandedMatrix)."""
    if isinstance(other, torch.Tensor):
        out_rows = A.shape[0]
        y = torch.empty((out_rows,) + tuple(other.shape[1:]), dtype=other.dtype, device=other.device)
        return mul_(y, A, other)
    if isinstance(A, BandedMatrix) and isinstance(other, BandedMatrix):
        return gbmm(A, other)
    from .special import structured_matmul
    from .tridiag import _Structured
    if isinstance(A, (BandedMatrix, _Structured)) and isinstance(other, (BandedMatrix, _Structured)):
        return structured_matmul(A, other)  # all-pairs products of test/test_special.jl:154-160
    raise LbmArgumentError(f"* is not defined for {type(A).__name__} and {type(other).__name__}")


def ldiv(A, b, out=None):
    """`A \\ b` for the structured kinds (Tridiagonal / SymTridiagonal / Bidiagonal): lbm_?tridiag_solve."""
    if hasattr(A, "solve"):
        return A.solve(b, out)
    raise LbmArgumentError(f"\\ is not defined for {type(A).__name__} on the B200 path")


def bandwidths(A):
    return tuple(A.bandwidths())


def blockbandwidths(A):
    return tuple(A.blockbandwidths())


def subblockbandwidths(A):
    return tuple(A.subblockbandwidths())
