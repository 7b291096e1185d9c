"""ctypes binding of liblbm_b200.so -- a 1:1 mirror of the `ccall`s in julia/LBMB200.jl.

The product has NO CPU fallback: if the shared library is missing, or no CUDA device is
present when a compute entry point is reached, this module raises -- it never routes
anywhere else.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading

from .errors import LbmArgumentError, LbmCudaError, LbmLibraryMissing

_HERE = os.path.dirname(os.path.abspath(__file__))
# LBM_B200_LIB: same override the Julia shim honours (benchmark A/B runs against another build of the library)
SO_PATH = os.environ.get("LBM_B200_LIB") or os.path.join(_HERE, "liblbm_b200.so")
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "lbm_b200.h"))

_lib = None
_lock = threading.Lock()

i64 = C.c_int64
u64 = C.c_uint64
vp = C.c_void_p
f64 = C.c_double
f32 = C.c_float
ch = C.c_char
sz = C.c_size_t


def _sigs():
    H = vp
    s = {
        "lbm_version": (C.c_int, []),
        "lbm_create": (C.c_int, [C.POINTER(vp), C.c_int]),
        "lbm_destroy": (C.c_int, [H]),
        "lbm_set_stream": (C.c_int, [H, vp]),
        "lbm_use_own_stream": (C.c_int, [H]),
        "lbm_sync": (C.c_int, [H]),
        "lbm_last_error_string": (C.c_char_p, [H]),
        "lbm_launch_count": (i64, [H]),
        "lbm_set_tuning": (C.c_int, [H, C.c_char_p, i64]),
        "lbm_get_tuning": (i64, [H, C.c_char_p]),
        "lbm_malloc": (C.c_int, [H, C.POINTER(vp), sz]),
        "lbm_free": (C.c_int, [H, vp]),
        "lbm_host_alloc": (C.c_int, [H, C.POINTER(vp), sz]),
        "lbm_host_free": (C.c_int, [H, vp]),
        "lbm_memcpy_h2d": (C.c_int, [H, vp, vp, sz]),
        "lbm_memcpy_d2h": (C.c_int, [H, vp, vp, sz]),
        "lbm_memcpy_d2d": (C.c_int, [H, vp, vp, sz]),
        "lbm_memset0": (C.c_int, [H, vp, sz]),
        "lbm_dfill_uniform": (C.c_int, [H, u64, i64, i64, vp]),
        "lbm_sfill_uniform": (C.c_int, [H, u64, i64, i64, vp]),
        "lbm_prod_bandwidths": (C.c_int, [i64, i64, C.c_int, C.POINTER(i64), C.POINTER(i64),
                                          C.POINTER(i64), C.POINTER(i64)]),
        "lbm_shard_plan": (C.c_int, [i64, i64, i64, C.c_int, C.c_int, C.POINTER(i64)]),
        "lbm_comm_unique_id": (C.c_int, [vp]),
        "lbm_comm_init": (C.c_int, [H, vp, C.c_int, C.c_int]),
        "lbm_comm_destroy": (C.c_int, [H]),
        "lbm_peer_alloc": (C.c_int, [H, sz, vp]),
        "lbm_peer_connect": (C.c_int, [H, vp, vp]),
        "lbm_peer_disconnect": (C.c_int, [H]),
        "lbm_peer_active": (C.c_int, [H]),
        "lbm_halo_exchange": (C.c_int, [H, i64, i64, i64, vp, C.c_int]),
    }
    M, S, V = vp, vp, vp      # lbm_multi_t, lbm_shard_t, lbm_svec_t
    s.update({
        "lbm_create_multi": (C.c_int, [C.POINTER(vp), C.c_int, C.POINTER(C.c_int)]),
        "lbm_destroy_multi": (C.c_int, [M]),
        "lbm_multi_ndev": (C.c_int, [M]),
        "lbm_multi_handle": (vp, [M, C.c_int]),
        "lbm_multi_sync": (C.c_int, [M]),
        "lbm_multi_last_error_string": (C.c_char_p, [M]),
        "lbm_shard_create": (C.c_int, [M, C.c_int, i64, i64, i64, C.POINTER(vp)]),
        "lbm_shard_destroy": (C.c_int, [S]),
        "lbm_shard_distribute": (C.c_int, [S, vp, i64]),
        "lbm_shard_fill_uniform": (C.c_int, [S, u64]),
        "lbm_shard_slab": (C.c_int, [S, C.c_int, C.POINTER(vp), C.POINTER(i64)]),
        "lbm_svec_create": (C.c_int, [S, C.POINTER(vp)]),
        "lbm_svec_destroy": (C.c_int, [V]),
        "lbm_svec_distribute": (C.c_int, [V, vp]),
        "lbm_svec_gather": (C.c_int, [V, vp]),
        "lbm_svec_fill_uniform": (C.c_int, [V, u64]),
        "lbm_svec_buffer": (C.c_int, [V, C.c_int, C.POINTER(vp), C.POINTER(vp)]),
    })
    for p, ft in (("d", f64), ("s", f32)):
        s[f"lbm_{p}gbmv_sharded_all"] = (C.c_int, [S, ft, V, ft, V])
        s[f"lbm_{p}gbmv_chain_sharded_all"] = (C.c_int, [C.c_int, C.POINTER(vp), ft, V, ft, V])
        s[f"lbm_{p}gbmv"] = (C.c_int, [H, ch, i64, i64, i64, i64, ft, vp, i64, vp, ft, vp])
        s[f"lbm_{p}gbmv_chain"] = (C.c_int, [H, C.c_int, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64),
                                            C.POINTER(vp), C.POINTER(i64), ft, vp, ft, vp])
        s[f"lbm_{p}gbmv_multi"] = (C.c_int, [H, ch, i64, i64, i64, i64, ft, vp, i64, vp, i64, ft, vp, i64, i64])
        s[f"lbm_{p}gbmv2"] = (C.c_int, [H, i64, i64, ft, i64, i64, vp, i64, ft, i64, i64, vp, i64, vp, ft, vp])
        s[f"lbm_{p}tridiag_dot"] = (C.c_int, [H, i64, vp, vp, vp, vp, vp, vp])
        s[f"lbm_{p}tridiag_sum"] = (C.c_int, [H, i64, vp, vp, vp, C.c_int, vp])
        s[f"lbm_{p}tridiag_axpby"] = (C.c_int, [H, i64, ft, vp, vp, vp, ft, vp, vp, vp, ft, vp, vp, vp])
        s[f"lbm_{p}tridiag_mm"] = (C.c_int, [H, i64, vp, vp, vp, vp, vp, vp, vp, i64])
        s[f"lbm_{p}tridiag_pack"] = (C.c_int, [H, i64, vp, vp, vp, i64, i64, vp, i64])
        s[f"lbm_{p}gbmv_host"] = (C.c_int, [H, ch, i64, i64, i64, i64, ft, vp, i64, vp, ft, vp])
        s[f"lbm_{p}gbmv_sharded"] = (C.c_int, [H, i64, i64, i64, ft, vp, i64, vp, ft, vp, C.c_int])
        s[f"lbm_{p}gbmv_chain_sharded"] = (C.c_int, [H, C.c_int, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(vp), C.POINTER(i64),
                                                    ft, vp, ft, vp])
        s[f"lbm_{p}tridiag_mv_sharded"] = (C.c_int, [H, i64, vp, vp, vp, ft, vp, ft, vp, C.c_int])
        s[f"lbm_{p}tridiag_solve"] = (C.c_int, [H, i64, vp, vp, vp, vp, i64, i64, vp, i64, vp])
        s[f"lbm_{p}bidiag_solve"] = (C.c_int, [H, ch, i64, vp, vp, vp, i64, i64, vp, i64, vp])
        s[f"lbm_{p}tridiag_mv"] = (C.c_int, [H, i64, vp, vp, vp, ft, vp, i64, i64, ft, vp, i64])
        s[f"lbm_{p}bidiag_mv"] = (C.c_int, [H, ch, i64, vp, vp, ft, vp, i64, i64, ft, vp, i64])
        s[f"lbm_{p}tridiag_mv_rows"] = (C.c_int, [H, i64, vp, vp, vp, ft, vp, ft, vp, i64, i64])
        s[f"lbm_{p}gbmm"] = (C.c_int, [H, i64, i64, i64, ft, i64, i64, vp, i64, i64, i64, vp, i64, ft,
                                       i64, i64, vp, i64])
        s[f"lbm_{p}gbmm_chain3_batched"] = (C.c_int, [H, i64, i64, C.POINTER(i64), C.POINTER(i64),
                                                      vp, i64, vp, i64, vp, i64, vp, i64])
        s[f"lbm_{p}kronsum2d_mv"] = (C.c_int, [H, i64, i64, i64, i64, vp, i64, i64, i64, vp, i64, ft,
                                               vp, ft, vp])
        s[f"lbm_{p}kronsum2d_mv_sharded"] = (C.c_int, [H, i64, i64, i64, i64, vp, i64, i64, i64, vp, i64, ft,
                                                       vp, ft, vp, C.c_int])
    s["lbm_diagtrav_length"] = (i64, [C.c_int, i64, i64])
    for p, ft in (("d", f64), ("s", f32)):
        s[f"lbm_{p}diagtrav"] = (C.c_int, [H, C.c_int, i64, i64, vp, vp])
        s[f"lbm_{p}invdiagtrav"] = (C.c_int, [H, C.c_int, i64, i64, vp, vp])
        s[f"lbm_{p}zero_bottomright"] = (C.c_int, [H, C.c_int, i64, i64, vp])
        s[f"lbm_{p}kron2d_mv"] = (C.c_int, [H, i64, i64, i64, i64, vp, i64, i64, i64, vp, i64, ft, vp, ft, vp])
        s[f"lbm_{p}kron3d_mv"] = (C.c_int, [H, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64, i64, vp, i64, ft, vp, ft, vp])
    return s


SIGNATURES = _sigs()


def header_symbols(path: str = HEADER_PATH):
    """Every function name include/lbm_b200.h declares (used by the ABI test)."""
    with open(path, encoding="utf-8") as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", src)))


def lib() -> C.CDLL:
    """Load the shared library (no GPU needed just to load it)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(SO_PATH):
                raise LbmLibraryMissing(
                    f"{SO_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU fallback)")
            L = C.CDLL(SO_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


class Handle:
    """One per device per host thread (`lbm_create`).  Kernels enqueue on `stream`
    (defaults to torch's current stream at call time when torch drives the device)."""

    def __init__(self, device: int = 0):
        self._h = vp()
        L = lib()
        rc = L.lbm_create(C.byref(self._h), int(device))
        if rc != 0:
            msg = L.lbm_last_error_string(None).decode()
            raise LbmCudaError(rc, f"lbm_create(device={device}) failed: {msg} "
                                   "(liblbm_b200 has no CPU fallback)")
        self.device = int(device)
        self.comm_world = 1  # > 1 once init_comm() has created the library's NCCL communicator

    @property
    def ptr(self):
        return self._h

    def check(self, rc: int, what: str = ""):
        if rc == 0:
            return
        msg = lib().lbm_last_error_string(self._h).decode()
        if rc < 0:
            # LAPACK-info style (-k = k-th argument).  Shape errors never get this far: the host
            # layer throws DimensionMismatch BEFORE the call, as the reference does
            # (src/bidiag.jl:363-377, test/test_tridiag.jl:272-279).
            raise LbmArgumentError(f"{what}: {msg}")
        raise LbmCudaError(rc, f"{what}: {msg}")

    def set_stream(self, cuda_stream_ptr):
        """`cuda_stream_ptr` is a cudaStream_t as an integer; 0 = CUDA's legacy default stream
        (which is what torch's default stream is)."""
        self.check(lib().lbm_set_stream(self._h, vp(int(cuda_stream_ptr) or None)))
        self._bound_stream = int(cuda_stream_ptr)

    def use_own_stream(self):
        self.check(lib().lbm_use_own_stream(self._h))
        self._bound_stream = None

    def use_torch_stream(self):
        import torch
        try:    # raw handle of torch's current stream without building a Stream object (small applies are host-bound)
            s = int(torch._C._cuda_getCurrentRawStream(self.device))
        except Exception:
            s = int(torch.cuda.current_stream(self.device).cuda_stream)
        if getattr(self, "_bound_stream", None) != s:   # small-n calls are host-bound: skip the FFI call when nothing changed
            self.set_stream(s)

    def sync(self):
        self.check(lib().lbm_sync(self._h), "lbm_sync")

    def launch_count(self) -> int:
        return int(lib().lbm_launch_count(self._h))

    def set_tuning(self, **kw):
        for k, v in kw.items():
            self.check(lib().lbm_set_tuning(self._h, k.encode(), int(v)), f"lbm_set_tuning({k})")

    def get_tuning(self, key: str) -> int:
        return int(lib().lbm_get_tuning(self._h, key.encode()))

    def close(self):
        if self._h:
            lib().lbm_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handles = {}


def handle_for(device: int) -> Handle:
    """Process-wide handle cache keyed by (thread, device)."""
    key = (threading.get_ident(), int(device))
    h = _handles.get(key)
    if h is None:
        h = Handle(device)
        _handles[key] = h
    return h


def handle_of(t) -> Handle:
    """Handle for the device a torch tensor lives on, bound to torch's current stream."""
    if not t.is_cuda:
        raise LbmCudaError(0, "operands must be CUDA-resident tensors: liblbm_b200 has no CPU path")
    h = handle_for(t.device.index if t.device.index is not None else 0)
    h.use_torch_stream()
    return h
