"""`+`, `-`, scalar `*` and `*` among the structured kinds (Diagonal, Bidiagonal, Tridiagonal,
SymTridiagonal, UniformScaling) with GPU-resident diagonals.

Host-side mirror of /root/reference/src/special.jl:75-222 (mixed-kind `+`/`-`, UniformScaling),
src/tridiag.jl:206-211, :570-575 and src/bidiag.jl:330-351 (same-kind `+`/`-`, unary `-`, scalar `*`).
Every method of the reference is "add/subtract the matching diagonals and wrap them in the smallest
kind that holds both operands"; here that is ONE launch of lbm_?tridiag_axpby over the (up to three)
result diagonals.  Like the reference, a result diagonal that is just one operand's diagonal
(`Bidiagonal(newdv, A.ev, A.uplo)`, special.jl:85-88) re-uses that operand's array.

Products of two structured operands (test/test_special.jl:154-160, test/test_bidiag.jl:262-267) go
through lbm_?tridiag_mm (the product band written in one pass; a BandedMatrix operand goes through `BandedMatrix(A)` =
lbm_?tridiag_pack and lbm_?gbmm), giving the banded result the reference's
ArrayLayouts route produces (bandwidths = sums).
"""
from __future__ import annotations

import torch

from . import _lib
from .banded import BandedMatrix, _ft, _pfx, _ptr, gbmm
from .errors import DimensionMismatch, LbmArgumentError
from .tridiag import Bidiagonal, SymTridiagonal, Tridiagonal, _Structured, _tridiag_mv_


class UniformScaling:
    """`λ*I` (LinearAlgebra.UniformScaling): only its `+`/`-` with structured kinds is on the path
    (src/special.jl:180-222)."""

    def __init__(self, lam):
        self.lam = float(lam)

    def __neg__(self):
        return UniformScaling(-self.lam)

    def __add__(self, other):
        return add(self, other)

    def __sub__(self, other):
        return sub(self, other)


class Diagonal(_Structured):
    """`Diagonal(diag)` (LinearAlgebra's type, an operand of src/special.jl:75-142)."""

    def __init__(self, diag: torch.Tensor):
        self.diag = diag.contiguous()
        self.n = int(diag.shape[0])

    @property
    def d(self):
        return self.diag

    def bandwidths(self):
        return (0, 0)

    def diagonaldata(self):
        return self.diag

    @property
    def T(self):
        return self

    def to_dense(self):
        return torch.diag(self.diag)

    def _diagonals(self):
        return (None, self.diag, None)

    def mul_(self, y, x, alpha=1.0, beta=0.0):
        return _tridiag_mv_(y, self.n, None, self.diag, None, x, alpha, beta)


def _kind(A):
    if isinstance(A, Diagonal):
        return "D"
    if isinstance(A, Bidiagonal):
        return "BU" if A.uplo == "U" else "BL"
    if isinstance(A, SymTridiagonal):
        return "S"
    if isinstance(A, Tridiagonal):
        return "T"
    raise LbmArgumentError(f"{type(A).__name__} is not a structured kind of src/special.jl")


def _join(ka, kb):
    """Smallest kind holding both operands -- the return types of src/special.jl:75-178, src/bidiag.jl:330-346."""
    if ka == "D":
        return kb
    if kb == "D":
        return ka
    return ka if ka == kb else "T"


def _three(A):
    """(dl, d, du) views with exactly n-1 / n / n-1 entries; None where the kind has no such diagonal."""
    dl, d, du = A._diagonals()
    m = max(A.n - 1, 0)
    cut = lambda v: None if v is None else v[:m]
    return cut(dl), d, cut(du)


def _wrap(kind, dl, d, du):
    if kind == "D":
        return Diagonal(d)
    if kind == "BU":
        return Bidiagonal(d, du, "U")
    if kind == "BL":
        return Bidiagonal(d, dl, "L")
    if kind == "S":
        return SymTridiagonal(d, dl)
    return Tridiagonal(dl, d, du)


def _check_sym_ev(A):
    # Julia's `A.dl + B.ev` / `A.ev + B.ev` throws for an `ev` of length n (allowed by src/tridiag.jl:11)
    # meeting one of length n-1; the broadcast shapes must agree.
    if isinstance(A, SymTridiagonal) and A.ev.shape[0] != max(A.n - 1, 0):
        raise DimensionMismatch(
            f"dimensions must match: a has dims ({A.ev.shape[0]},), b has dims ({max(A.n - 1, 0)},)")


def axpby(alpha, A, beta, B, gamma=0.0):
    """`alpha*A + beta*B + gamma*I` on diagonals; `A` or `B` may be None.  One launch."""
    ops = [X for X in (A, B) if X is not None]
    if not ops:
        raise LbmArgumentError("axpby needs at least one structured operand")
    if A is not None and B is not None:
        if A.n != B.n:
            raise DimensionMismatch(f"dimensions must match: a has dims {A.shape}, b has dims {B.shape}")
        if A.dtype != B.dtype or A.device != B.device:
            raise LbmArgumentError("operands must share element type and device")
        if not (isinstance(A, SymTridiagonal) and isinstance(B, SymTridiagonal) and A.ev.shape == B.ev.shape):
            _check_sym_ev(A)
            _check_sym_ev(B)
    n, ref = ops[0].n, ops[0].d
    kind = _kind(ops[0]) if len(ops) == 1 else _join(_kind(A), _kind(B))
    a3 = _three(A) if A is not None else (None, None, None)
    b3 = _three(B) if B is not None else (None, None, None)
    want = {"D": (0, 1, 0), "BU": (0, 1, 1), "BL": (1, 1, 0), "S": (1, 1, 0), "T": (1, 1, 1)}[kind]
    coef = (alpha, beta)
    outs, ain, bin_, reuse = [None] * 3, [None] * 3, [None] * 3, [None] * 3
    for k in range(3):
        if not want[k]:
            continue
        a, b = a3[k], b3[k]
        g = gamma if k == 1 else 0.0
        have = [(c, v) for c, v in ((coef[0], a), (coef[1], b)) if v is not None]
        length = n if k == 1 else max(n - 1, 0)
        if len(have) == 1 and have[0][0] == 1.0 and g == 0.0:
            reuse[k] = have[0][1]          # the reference hands the operand's own array on
            continue
        outs[k] = torch.empty(length, dtype=ref.dtype, device=ref.device)
        if not have:
            outs[k].zero_()                # unreachable for the reference's methods; kept total
            reuse[k], outs[k] = outs[k], None
            continue
        ain[k], bin_[k] = a, b
    if n > 0 and any(o is not None for o in outs):
        h = _lib.handle_of(ref)
        fn = getattr(_lib.lib(), f"lbm_{_pfx(ref.dtype)}tridiag_axpby")
        ft = _ft(ref.dtype)
        # a diagonal whose output is not wanted must not be read either
        pa = [ain[k] if outs[k] is not None else None for k in range(3)]
        pb = [bin_[k] if outs[k] is not None else None for k in range(3)]
        h.check(fn(h.ptr, n, ft(alpha), _ptr(pa[0]), _ptr(pa[1]), _ptr(pa[2]), ft(beta), _ptr(pb[0]), _ptr(pb[1]),
                   _ptr(pb[2]), ft(gamma), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2])), "lbm_tridiag_axpby")
    res = [outs[k] if outs[k] is not None else reuse[k] for k in range(3)]
    return _wrap(kind, *res)


def add(A, B):
    """`A + B` (src/special.jl:75-178 mixed kinds, :180-206 UniformScaling; same kinds src/tridiag.jl:206,:571,
    src/bidiag.jl:330-337)."""
    if isinstance(B, UniformScaling):
        return axpby(1.0, A, 0.0, None, B.lam)
    if isinstance(A, UniformScaling):
        return axpby(0.0, None, 1.0, B, A.lam)
    return axpby(1.0, A, 1.0, B)


def sub(A, B):
    """`A - B` (src/special.jl:80-178, :208-222; same kinds src/tridiag.jl:207,:572, src/bidiag.jl:339-346)."""
    if isinstance(B, UniformScaling):
        return axpby(1.0, A, 0.0, None, -B.lam)
    if isinstance(A, UniformScaling):
        return axpby(0.0, None, -1.0, B, A.lam)
    return axpby(1.0, A, -1.0, B)


def neg(A):
    """`-A` (src/tridiag.jl:208, :570; src/bidiag.jl:348)."""
    return axpby(-1.0, A, 0.0, None)


def scale(A, c):
    """`A*c`, `c*A` for a number c (src/tridiag.jl:209-210, :573-574; src/bidiag.jl:349-350)."""
    return axpby(float(c), A, 0.0, None)


def to_banded(A, l=None, u=None) -> BandedMatrix:
    """`BandedMatrix(A, (l,u))` of a structured operand (defaults: its own bandwidths)."""
    bl, bu = A.bandwidths()
    l = bl if l is None else int(l)
    u = bu if u is None else int(u)
    if l < bl or u < bu:
        raise LbmArgumentError(f"bandwidths ({l},{u}) cannot hold a matrix with bandwidths ({bl},{bu})")
    dl, d, du = A._diagonals()
    data = torch.empty((A.n, l + u + 1), dtype=d.dtype, device=d.device)
    if A.n:
        h = _lib.handle_of(d)
        fn = getattr(_lib.lib(), f"lbm_{_pfx(d.dtype)}tridiag_pack")
        h.check(fn(h.ptr, A.n, _ptr(dl), _ptr(d), _ptr(du), l, u, _ptr(data), l + u + 1), "lbm_tridiag_pack")
    return BandedMatrix(data, A.n, l, u)


def structured_matmul(A, B) -> BandedMatrix:
    """`A*B` for two structured / banded operands -> BandedMatrix with bandwidths (la+lb, ua+ub)."""
    if not isinstance(A, BandedMatrix) and not isinstance(B, BandedMatrix):
        if A.n != B.n:
            raise DimensionMismatch(
                f"second entry of size(A)={A.shape} and first entry of size(B) = {B.shape} must match.")
        if A.dtype != B.dtype or A.device != B.device:
            raise LbmArgumentError("operands must share element type and device")
        (la, ua), (lb, ub) = A.bandwidths(), B.bandwidths()
        rows = la + lb + ua + ub + 1
        out = torch.empty((A.n, rows), dtype=A.dtype, device=A.device)
        if A.n:
            adl, ad, adu = A._diagonals()
            bdl, bd, bdu = B._diagonals()
            if A.n == 1:   # no off-diagonal entries: keep the band shape, pass the (empty) diagonals as present
                z = torch.zeros(1, dtype=A.dtype, device=A.device)
                adl, adu = (z if la else None), (z if ua else None)
                bdl, bdu = (z if lb else None), (z if ub else None)
            h = _lib.handle_of(ad)
            fn = getattr(_lib.lib(), f"lbm_{_pfx(ad.dtype)}tridiag_mm")
            h.check(fn(h.ptr, A.n, _ptr(adl), _ptr(ad), _ptr(adu), _ptr(bdl), _ptr(bd), _ptr(bdu), _ptr(out), rows),
                    "lbm_tridiag_mm")
        return BandedMatrix(out, A.n, la + lb, ua + ub)
    Ab = A if isinstance(A, BandedMatrix) else to_banded(A)
    Bb = B if isinstance(B, BandedMatrix) else to_banded(B)
    return gbmm(Ab, Bb)


def _install_operators():
    def _add(self, other):
        return add(self, other)

    def _radd(self, other):
        return add(other, self)

    def _sub(self, other):
        return sub(self, other)

    def _rsub(self, other):
        return sub(other, self)

    def _mul(self, c):
        if isinstance(c, (int, float)):
            return scale(self, c)
        return NotImplemented

    _Structured.__add__ = _add
    _Structured.__radd__ = _radd
    _Structured.__sub__ = _sub
    _Structured.__rsub__ = _rsub
    _Structured.__neg__ = neg
    _Structured.__mul__ = _mul
    _Structured.__rmul__ = _mul


_install_operators()
