"""`BlockKron` / `KronTrav` structure and the never-materialised Kronecker applies.

  * `BlockKron(A, B)` ([up] BlockArrays, re-exported by the reference at
    /root/reference/src/LazyBandedMatrices.jl:41): blocks `A[K,J]*B`;
    blockbandwidths = bandwidths(A), subblockbandwidths = bandwidths(B)
    (pinned by test/test_blockkron.jl:94-107).  `K @ x` = vec(B X A^T) -> lbm_dkron2d_mv.
  * `BroadcastArray(+, BlockKron(A, Eye), BlockKron(Eye, B))` -- the 2-D Laplacian of
    test/test_blockkron.jl:296-306 -- applies as ONE fused kernel, lbm_?kronsum2d_mv.
  * `KronTrav(A, B, ...)` structure restated from src/blockkron.jl:353-362.
"""
from __future__ import annotations

import torch

from . import _lib
from .banded import BandedMatrix, _ft, _pfx, _ptr
from .blockconcat import BlockBandedSpec
from .errors import DimensionMismatch, LbmArgumentError


class Eye:
    """`Eye(n)`: identity with bandwidths (0,0) ([up] FillArrays)."""

    def __init__(self, n: int):
        self.n = int(n)

    @property
    def shape(self):
        return (self.n, self.n)

    def bandwidths(self):
        return (0, 0)

    def as_banded(self, dtype, device) -> BandedMatrix:
        return BandedMatrix(torch.ones((self.n, 1), dtype=dtype, device=device), self.n, 0, 0)


def _bw(a):
    return tuple(a.bandwidths())


class BlockKron:
    def __init__(self, *args):
        if len(args) < 2:
            raise LbmArgumentError("BlockKron needs at least two factors")
        self.args = args

    @property
    def shape(self):
        m = n = 1
        for a in self.args:
            m *= a.shape[0]
            n *= a.shape[1]
        return (m, n)

    def blockbandwidths(self):
        return _bw(self.args[0])  # [up] BlockBandedMatrices: bandwidths(K.args[1])

    def subblockbandwidths(self):
        # [up]: bandwidths(K.args[2]) -- exact for 2 factors; for 3 factors this is the upstream
        # defect recorded @test_broken at test/test_blockkron.jl:109-113 (not replicated further).
        return _bw(self.args[1])

    def isblockbanded(self):
        return True

    def isbandedblockbanded(self):
        return True

    def blocksize(self):
        return tuple(self.args[0].shape)

    def block_spec(self):
        return BlockBandedSpec(self.blocksize(), self.blockbandwidths(), self.subblockbandwidths())

    # ---- apply ----------------------------------------------------------------------
    def mul_(self, y, x, alpha=1.0, beta=0.0):
        if len(self.args) != 2:
            raise LbmArgumentError("only 2-factor BlockKron applies are in scope")
        A, B = self.args
        n1, n2 = A.shape[1], B.shape[1]
        if A.shape[0] != A.shape[1] or B.shape[0] != B.shape[1]:
            raise LbmArgumentError("BlockKron apply requires square factors")
        if x.shape[0] != n1 * n2:
            raise DimensionMismatch(
                f"second entry of size(A)={self.shape} and first entry of size(B) = {tuple(x.shape)} must match.")
        if y.shape[0] != n1 * n2:
            raise DimensionMismatch(f"sizes size(A)={self.shape} and size(C) = {tuple(y.shape)} must match at first entry.")
        if x.dtype != torch.float64:
            raise LbmArgumentError("lbm_dkron2d_mv is Float64")
        Ab = A.as_banded(x.dtype, x.device) if isinstance(A, Eye) else A
        Bb = B.as_banded(x.dtype, x.device) if isinstance(B, Eye) else B
        h = _lib.handle_of(x)
        h.check(_lib.lib().lbm_dkron2d_mv(h.ptr, n1, n2, Ab.l, Ab.u, _ptr(Ab.data), Ab.lda, Bb.l, Bb.u,
                                          _ptr(Bb.data), Bb.lda, float(alpha), _ptr(x), float(beta), _ptr(y)),
                "lbm_dkron2d_mv")
        return y

    def __matmul__(self, x):
        y = torch.empty_like(x)
        return self.mul_(y, x)


class KronSum:
    """`BroadcastArray(+, BlockKron(A, Eye(n2)), BlockKron(Eye(n1), B))`: kron(A,I) + kron(I,B).
    blockbandwidths/subblockbandwidths = max. over the two terms (a broadcast sum keeps the
    union band; pinned (1,1)/(1,1) at test/test_blockkron.jl:305-306)."""

    def __init__(self, A: BandedMatrix, B: BandedMatrix):
        if A.shape[0] != A.shape[1] or B.shape[0] != B.shape[1]:
            raise LbmArgumentError("KronSum requires square factors")
        self.A, self.B = A, B

    @classmethod
    def from_broadcast(cls, t1: BlockKron, t2: BlockKron):
        a1, b1 = t1.args
        a2, b2 = t2.args
        if isinstance(b1, Eye) and isinstance(a2, Eye) and b1.n == b2.shape[0] and a2.n == a1.shape[0]:
            return cls(a1, b2)
        if isinstance(a1, Eye) and isinstance(b2, Eye) and b2.n == b1.shape[0] and a1.n == a2.shape[0]:
            return cls(a2, b1)
        raise LbmArgumentError("expected BlockKron(A, Eye) + BlockKron(Eye, B)")

    @property
    def shape(self):
        n = self.A.shape[0] * self.B.shape[0]
        return (n, n)

    def terms(self):
        return (BlockKron(self.A, Eye(self.B.shape[0])), BlockKron(Eye(self.A.shape[0]), self.B))

    def blockbandwidths(self):
        t1, t2 = self.terms()
        a, b = t1.blockbandwidths(), t2.blockbandwidths()
        return (max(a[0], b[0]), max(a[1], b[1]))

    def subblockbandwidths(self):
        t1, t2 = self.terms()
        a, b = t1.subblockbandwidths(), t2.subblockbandwidths()
        return (max(a[0], b[0]), max(a[1], b[1]))

    def mul_(self, y, x, alpha=1.0, beta=0.0):
        n1, n2 = self.A.shape[0], self.B.shape[0]
        if x.shape[0] != n1 * n2:
            raise DimensionMismatch(
                f"second entry of size(A)={self.shape} and first entry of size(B) = {tuple(x.shape)} must match.")
        if y.shape[0] != n1 * n2:
            raise DimensionMismatch(f"sizes size(A)={self.shape} and size(C) = {tuple(y.shape)} must match at first entry.")
        h = _lib.handle_of(x)
        fn = getattr(_lib.lib(), f"lbm_{_pfx(x.dtype)}kronsum2d_mv")
        ft = _ft(x.dtype)
        A, B = self.A, self.B
        h.check(fn(h.ptr, n1, n2, A.l, A.u, _ptr(A.data), A.lda, B.l, B.u, _ptr(B.data), B.lda, ft(alpha),
                   _ptr(x), ft(beta), _ptr(y)), "lbm_kronsum2d_mv")
        return y

    def __matmul__(self, x):
        y = torch.empty_like(x)
        return self.mul_(y, x)


class KronTrav:
    """Structure of `KronTrav(A, B[, C])` (Kronecker product in DiagTrav ordering)."""

    def __init__(self, *args):
        self.args = args

    def blockbandwidths(self):
        # _krontrav_blockbandwidths: broadcast(+, map(bandwidths, A)...)  (src/blockkron.jl:355,358)
        l = sum(_bw(a)[0] for a in self.args)
        u = sum(_bw(a)[1] for a in self.args)
        return (l, u)

    def subblockbandwidths(self):
        # src/blockkron.jl:353-354,357
        if len(self.args) == 2:
            return _bw(self.args[0])
        if len(self.args) == 3:
            return KronTrav(self.args[0], self.args[1]).blockbandwidths()
        raise LbmArgumentError("subblockbandwidths(KronTrav) is defined for 2 or 3 factors")

    def isblockbanded(self):
        return True  # all(isbanded, args): every operand here is banded (src/blockkron.jl:361)

    def isbandedblockbanded(self):
        return self.isblockbanded() and len(self.args) == 2  # src/blockkron.jl:362
