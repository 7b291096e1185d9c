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
        Ab = A.as_banded(x.dtype, x.device) if isinstance(A, Eye) else A
        Bb = B.as_banded(x.dtype, x.device) if isinstance(B, Eye) else B
        h = _lib.handle_of(x)
        fn = getattr(_lib.lib(), f"lbm_{_pfx(x.dtype)}kron2d_mv")
        ft = _ft(x.dtype)
        h.check(fn(h.ptr, n1, n2, Ab.l, Ab.u, _ptr(Ab.data), Ab.lda, Bb.l, Bb.u,
                   _ptr(Bb.data), Bb.lda, ft(alpha), _ptr(x), ft(beta), _ptr(y)), "lbm_kron2d_mv")
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


def _diagtrav_index(shape):
    """Flat (column-major) indices of an array's entries in DiagTrav order: anti-diagonal blocks
    K = 1.., inside a block k descending (/root/reference/src/blockkron.jl:14-24, :101-133; 3-D order
    pinned by test/test_blockkron.jl:53-57).  Host integer logic only."""
    if len(shape) == 2:
        m, n = shape
        out = []
        for K in range(1, max(m, n) + 1):
            for j in range(1, K + 1):
                k = K + 1 - j
                if k <= m and j <= n:
                    out.append((k - 1) + m * (j - 1))
        return out
    if len(shape) == 3:
        n = shape[0]
        if not (shape[1] == n and shape[2] == n):
            raise LbmArgumentError("3-D DiagTrav needs a cubic array")
        out = []
        for K in range(1, n + 1):
            for k in range(K, 0, -1):
                for j in range(K + 1 - k, 0, -1):
                    l = K + 2 - k - j
                    out.append((k - 1) + n * (j - 1) + n * n * (l - 1))
        return out
    raise LbmArgumentError("DiagTrav supports 2-D and 3-D arrays")


class DiagTrav:
    """`DiagTrav(X)`: a matrix / cubic tensor viewed as a block vector by anti-diagonal traversal.
    The constructor zeroes the bottom-right part (zero_bottomright, src/blockkron.jl:27-64).

    Storage is the column-major vec of the array (k fastest) -- what the Julia side holds natively and what the Kronecker
    kernels take -- so `KronTrav.mul_diagtrav` runs kernel to kernel with no re-layout in between; the DiagTrav ordering,
    its inverse and zero_bottomright are device kernels (lbm_?diagtrav / lbm_?invdiagtrav / lbm_?zero_bottomright)."""

    def __init__(self, array: torch.Tensor = None, *, _flat: torch.Tensor = None, _shape=None):
        if _flat is None:
            if array.dim() not in (2, 3):
                raise LbmArgumentError("DiagTrav supports 2-D and 3-D arrays")
            if array.dim() == 3 and not (array.shape[0] == array.shape[1] == array.shape[2]):
                raise LbmArgumentError("3-D DiagTrav needs a cubic array")
            _shape = tuple(int(v) for v in array.shape)
            # torch indexes [k, j(, l)] row-major; the one-time conversion to the column-major vec happens HERE, outside the apply
            _flat = array.permute(*reversed(range(array.dim()))).contiguous().view(-1).clone()
        self.shape = tuple(_shape)
        self._flat = _flat
        if _flat.numel():
            h = _lib.handle_of(_flat)
            m = self.shape[0]
            n = self.shape[1]
            fn = getattr(_lib.lib(), f"lbm_{_pfx(_flat.dtype)}zero_bottomright")
            h.check(fn(h.ptr, len(self.shape), m, n, _ptr(_flat)), "lbm_zero_bottomright")

    @property
    def array(self) -> torch.Tensor:
        """The array indexed [k, j(, l)] (a VIEW of the column-major storage)."""
        return self._flat.view(*reversed(self.shape)).permute(*reversed(range(len(self.shape))))

    def colmajor(self) -> torch.Tensor:
        """vec(array) with k fastest (the layout the Kronecker kernels take)."""
        return self._flat

    def __len__(self):
        return int(_lib.lib().lbm_diagtrav_length(len(self.shape), self.shape[0], self.shape[1]))

    def vector(self) -> torch.Tensor:
        """The block vector (what `Vector(DiagTrav(X))` returns in the reference): lbm_?diagtrav."""
        out = torch.empty(len(self), dtype=self._flat.dtype, device=self._flat.device)
        if out.numel():
            h = _lib.handle_of(self._flat)
            fn = getattr(_lib.lib(), f"lbm_{_pfx(self._flat.dtype)}diagtrav")
            h.check(fn(h.ptr, len(self.shape), self.shape[0], self.shape[1], _ptr(self._flat), _ptr(out)), "lbm_diagtrav")
        return out

    @classmethod
    def from_vector(cls, v: torch.Tensor, shape):
        """`InvDiagTrav`: the array whose DiagTrav is the block vector `v` (bottom-right zero): lbm_?invdiagtrav."""
        shape = tuple(int(s) for s in shape)
        want = int(_lib.lib().lbm_diagtrav_length(len(shape), shape[0], shape[1]))
        if v.numel() != want:
            raise DimensionMismatch(f"block vector has {v.numel()} entries, DiagTrav of a {shape} array has {want}")
        total = 1
        for d in shape:
            total *= d
        flat = torch.empty(total, dtype=v.dtype, device=v.device)
        if total:
            h = _lib.handle_of(v)
            fn = getattr(_lib.lib(), f"lbm_{_pfx(v.dtype)}invdiagtrav")
            h.check(fn(h.ptr, len(shape), shape[0], shape[1], _ptr(v.contiguous()), _ptr(flat)), "lbm_invdiagtrav")
        obj = cls.__new__(cls)
        obj.shape, obj._flat = shape, flat
        return obj

    @classmethod
    def from_colmajor(cls, v: torch.Tensor, shape):
        return cls(_flat=v, _shape=shape)


class KronTrav:
    """`KronTrav(A, B[, C])`: Kronecker product in DiagTrav ordering -- structure queries plus the
    never-materialised apply `K*DiagTrav(X) = DiagTrav(B*X*A')` (src/blockkron.jl:433-440) and its
    3-D version by mode products (:441-450)."""

    def __init__(self, *args):
        self.args = args

    def mul_diagtrav(self, X: "DiagTrav") -> "DiagTrav":
        """Kernel to kernel: lbm_?kron2d_mv / lbm_?kron3d_mv on the column-major storage, then the DiagTrav constructor's
        zero_bottomright kernel -- no eager tensor ops in between."""
        x = X.colmajor()
        h = _lib.handle_of(x)
        y = torch.empty_like(x)
        ft = _ft(x.dtype)
        bm = [a.as_banded(x.dtype, x.device) if isinstance(a, Eye) else a for a in self.args]
        if len(bm) == 2 and len(X.shape) == 2:
            A, B = bm
            n2, n1 = X.shape                        # X is size(B,2) x size(A,2)
            if (A.shape[1], B.shape[1]) != (n1, n2):
                raise DimensionMismatch(f"KronTrav factors {A.shape}, {B.shape} do not match DiagTrav array {tuple(X.shape)}")
            fn = getattr(_lib.lib(), f"lbm_{_pfx(x.dtype)}kron2d_mv")
            h.check(fn(h.ptr, n1, n2, A.l, A.u, _ptr(A.data), A.lda, B.l, B.u, _ptr(B.data), B.lda, ft(1.0), _ptr(x), ft(0.0),
                       _ptr(y)), "lbm_kron2d_mv")
            return DiagTrav.from_colmajor(y, (n2, n1))
        if len(bm) == 3 and len(X.shape) == 3:
            A, B, Cm = bm
            n = X.shape[0]
            if not all(f.shape == (n, n) for f in bm):
                raise DimensionMismatch("3-factor KronTrav apply needs n x n factors and an n^3 tensor")
            fn = getattr(_lib.lib(), f"lbm_{_pfx(x.dtype)}kron3d_mv")
            h.check(fn(h.ptr, n, A.l, A.u, _ptr(A.data), A.lda, B.l, B.u, _ptr(B.data), B.lda,
                       Cm.l, Cm.u, _ptr(Cm.data), Cm.lda, ft(1.0), _ptr(x), ft(0.0), _ptr(y)), "lbm_kron3d_mv")
            return DiagTrav.from_colmajor(y, (n, n, n))
        raise LbmArgumentError("KronTrav apply is defined for 2 factors x matrix or 3 factors x cubic tensor")

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
