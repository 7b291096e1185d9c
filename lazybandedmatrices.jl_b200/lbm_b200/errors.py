"""Exception types mirroring the reference's error behaviour."""


class DimensionMismatch(ValueError):
    """Julia's `DimensionMismatch` (thrown before any FFI call; messages follow
    /root/reference/src/bidiag.jl:363-377)."""


class LbmArgumentError(ValueError):
    """Julia's `ArgumentError` / a negative LAPACK-style `info` from the C ABI."""


class SingularException(ArithmeticError):
    """Julia's `SingularException`: a zero pivot in a factorisation / solve."""


class LbmCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[cuda {code}] {msg}")
        self.code = code


class LbmLibraryMissing(ImportError):
    """liblbm_b200.so is not built; the product has no CPU fallback."""
