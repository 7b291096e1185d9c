"""lbm_b200 -- host-side mirror of LazyBandedMatrices.jl's public API over liblbm_b200.so.

Same names as the reference exports/uses (src/LazyBandedMatrices.jl:41 and the types of
src/tridiag.jl, src/bidiag.jl): operands are torch CUDA tensors, every `mul!`/`materialize`
lands in a hand-written sm_100a kernel through the C ABI of include/lbm_b200.h.  There is no
CPU path.
"""
from .errors import DimensionMismatch, LbmArgumentError, LbmCudaError, LbmLibraryMissing, SingularException  # noqa: F401
from ._lib import Handle, handle_for, lib, SO_PATH  # noqa: F401
from .banded import BandedMatrix, TransposedBanded, brand, banded_from_diagonals, fill_uniform, gbmm, gbmv_  # noqa: F401
from .tridiag import Tridiagonal, SymTridiagonal, Bidiagonal  # noqa: F401
from .special import Diagonal, UniformScaling, add, sub, neg, scale, axpby, to_banded, structured_matmul  # noqa: F401
from .lazy import ApplyMatrix, BroadcastMatrix, chain3_batched  # noqa: F401
from .blockkron import BlockKron, KronSum, KronTrav, DiagTrav, Eye  # noqa: F401
from .blockconcat import BlockVcat, BlockHcat, BlockBroadcastArray, BlockBandedSpec, unitblocks, zeros_spec  # noqa: F401
from .linalg import mul_, matmul, ldiv, bandwidths, blockbandwidths, subblockbandwidths  # noqa: F401

__all__ = [
    "BandedMatrix", "brand", "banded_from_diagonals", "Tridiagonal", "SymTridiagonal", "Bidiagonal",
    "Diagonal", "UniformScaling", "add", "sub", "neg", "scale", "axpby", "to_banded", "structured_matmul",
    "ApplyMatrix", "BroadcastMatrix", "chain3_batched", "BlockKron", "KronSum", "KronTrav", "DiagTrav", "Eye", "BlockVcat", "BlockHcat",
    "BlockBroadcastArray", "BlockBandedSpec", "unitblocks", "zeros_spec", "mul_", "matmul", "bandwidths",
    "blockbandwidths", "subblockbandwidths", "ldiv", "SingularException", "DimensionMismatch", "LbmArgumentError", "LbmCudaError",
    "Handle", "handle_for", "lib", "fill_uniform", "gbmm", "gbmv_",
]
