"""Structure queries of the reference's lazy block containers (host side, pure integers).

`north_star`: the host keeps the package's public API and dispatch (BlockVcat/BlockHcat,
BlockBroadcastArray, blockbandwidths, subblockbandwidths); results must match the reference
bit-exactly.  Only the STRUCTURE algebra is on the hot path -- the data movement of lazy
concatenation has no flops and stays out of scope.

Restated from /root/reference/src/blockconcat.jl:364-395 and src/LazyBandedMatrices.jl:50-52.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence, Tuple

NEG = -(1 << 40)  # stands in for the "-infinity" bandwidths of an all-zero operand


@dataclass(frozen=True)
class BlockBandedSpec:
    """What the structure algebra needs to know about one block-banded operand."""
    blocksize: Tuple[int, int]               # (block rows, block cols)
    blockbandwidths: Tuple[int, int]
    subblockbandwidths: Tuple[int, int] = (0, 0)
    isblockbanded: bool = True


def zeros_spec(blocksize: Tuple[int, int]) -> BlockBandedSpec:
    """`Zeros(axes(a))`: no stored blocks at all."""
    return BlockBandedSpec(tuple(blocksize), (NEG, NEG), (NEG, NEG))


def _spec(a) -> BlockBandedSpec:
    if isinstance(a, BlockBandedSpec):
        return a
    if hasattr(a, "block_spec"):
        return a.block_spec()
    raise TypeError(f"{type(a).__name__} has no block structure")


def unitblocks(a) -> BlockBandedSpec:
    """`unitblocks(A)` (src/LazyBandedMatrices.jl:50-52): every entry is its own 1x1 block, so
    blockbandwidths == bandwidths(A) and subblockbandwidths == (0,0)
    (test/test_blockconcat.jl:409-412; test/test_misc.jl:59)."""
    m, n = a.shape
    l, u = a.bandwidths()
    return BlockBandedSpec((m, n), (l, u), (0, 0))


class BlockVcat:
    """Lazy vertical concatenation of block matrices; structure only."""

    def __init__(self, *arrays):
        self.arrays = tuple(_spec(a) for a in arrays)

    def blockbandwidths(self):
        # src/blockconcat.jl:368-371
        cs = [0]
        for a in self.arrays[:-1]:
            cs.append(cs[-1] + a.blocksize[0])
        l = max(c + a.blockbandwidths[0] for c, a in zip(cs, self.arrays))
        u = max(a.blockbandwidths[1] - c for c, a in zip(cs, self.arrays))
        return (l, u)

    def isblockbanded(self):
        return all(a.isblockbanded for a in self.arrays)  # src/blockconcat.jl:372

    def blocksize(self):
        return (sum(a.blocksize[0] for a in self.arrays), self.arrays[0].blocksize[1])

    def block_spec(self):
        return BlockBandedSpec(self.blocksize(), self.blockbandwidths(), (NEG, NEG), self.isblockbanded())


class BlockHcat:
    """Lazy horizontal concatenation of block matrices; structure only."""

    def __init__(self, *arrays):
        self.arrays = tuple(_spec(a) for a in arrays)

    def blockbandwidths(self):
        # src/blockconcat.jl:374-377
        cs = [0]
        for a in self.arrays[:-1]:
            cs.append(cs[-1] + a.blocksize[1])
        l = max(a.blockbandwidths[0] - c for c, a in zip(cs, self.arrays))
        u = max(a.blockbandwidths[1] + c for c, a in zip(cs, self.arrays))
        return (l, u)

    def isblockbanded(self):
        return all(a.isblockbanded for a in self.arrays)  # src/blockconcat.jl:378

    def blocksize(self):
        return (self.arrays[0].blocksize[0], sum(a.blocksize[1] for a in self.arrays))

    def block_spec(self):
        return BlockBandedSpec(self.blocksize(), self.blockbandwidths(), (NEG, NEG), self.isblockbanded())


def _max2(a: Sequence[int], b: Sequence[int]):
    return (max(a[0], b[0]), max(a[1], b[1]))


class BlockBroadcastArray:
    """`BlockBroadcastArray(f, args...)` for f in {hvcat, Diagonal}: block-wise interlacing.
    `BlockBroadcastArray("hvcat", p, a11, a12, ..., app)` / `BlockBroadcastArray("Diagonal", A, B)`."""

    def __init__(self, f: str, *args):
        if f not in ("hvcat", "Diagonal"):
            raise ValueError("structure queries are defined for f in {'hvcat', 'Diagonal'}")
        self.f = f
        if f == "hvcat":
            self.p = int(args[0])
            self.args = tuple(_spec(a) for a in args[1:])
        else:
            self.p = None
            self.args = tuple(_spec(a) for a in args)

    def blockbandwidths(self):
        # src/blockconcat.jl:364-365: max.(map(blockbandwidths, args)...)
        out = self.args[0].blockbandwidths
        for a in self.args[1:]:
            out = _max2(out, a.blockbandwidths)
        return out

    def subblockbandwidths(self):
        if self.f == "Diagonal":
            return (0, 0)  # src/blockconcat.jl:366
        # src/blockconcat.jl:380-395
        p = self.p
        s0 = self.args[0].subblockbandwidths
        ret = (p * s0[0], p * s0[1])
        shft = 0
        rw = 0
        for a in self.args[1:]:
            shft += 1
            if shft == p:
                rw += 1
                shft = -rw
            s = a.subblockbandwidths
            ret = _max2(ret, (p * s[0] - shft, p * s[1] + shft))
        return ret
