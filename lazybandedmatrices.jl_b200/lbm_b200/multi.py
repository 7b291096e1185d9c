"""Single-process multi-device mode: ONE host thread drives every GPU of the box.

Host-side mirror of what the Julia shim does with `lbm_create_multi` / `lbm_shard_*` / `lbm_svec_*` /
`lbm_?gbmv_sharded_all`: a `ShardedDevBandedMatrix` whose `mul!` is one C call that enqueues one kernel per device
(ghost exchange fused into it).  No torch.distributed, no NCCL, no rendezvous.  Operands are numpy host arrays in
BandedMatrices' own storage ((l+u+1) x n, Fortran order) -- exactly what `bandeddata(A)` is on the Julia side.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .errors import DimensionMismatch, LbmArgumentError, LbmCudaError

_PFX = {np.dtype(np.float64): "d", np.dtype(np.float32): "s"}


class MultiDevice:
    """`lbm_create_multi(ndev, devices)`."""

    def __init__(self, ndev: int, devices=None):
        self._m = C.c_void_p()
        L = _lib.lib()
        devs = None if devices is None else (C.c_int * ndev)(*devices)
        rc = L.lbm_create_multi(C.byref(self._m), int(ndev), devs)
        if rc != 0:
            raise LbmCudaError(rc, f"lbm_create_multi({ndev}) failed ({rc}): {L.lbm_last_error_string(None).decode()}")
        self.ndev = int(ndev)

    @property
    def ptr(self):
        return self._m

    def check(self, rc, what=""):
        if rc == 0:
            return
        msg = _lib.lib().lbm_multi_last_error_string(self._m).decode()
        if rc < 0:
            if "DimensionMismatch" in msg:
                raise DimensionMismatch(msg)
            raise LbmArgumentError(f"{what}: {msg}")
        raise LbmCudaError(rc, f"{what}: {msg}")

    def handle(self, r: int):
        """The per-device handle (an `lbm_handle_t`) for the `*_sharded` entry points."""
        return _lib.lib().lbm_multi_handle(self._m, int(r))

    def sync(self):
        self.check(_lib.lib().lbm_multi_sync(self._m), "lbm_multi_sync")

    def launch_count(self) -> int:
        L = _lib.lib()
        return sum(int(L.lbm_launch_count(C.c_void_p(L.lbm_multi_handle(self._m, r)))) for r in range(self.ndev))

    def set_tuning(self, **kw):
        L = _lib.lib()
        for r in range(self.ndev):
            for k, v in kw.items():
                rc = L.lbm_set_tuning(C.c_void_p(L.lbm_multi_handle(self._m, r)), k.encode(), int(v))
                if rc:
                    raise LbmArgumentError(f"lbm_set_tuning({k})")

    def close(self):
        if self._m:
            _lib.lib().lbm_destroy_multi(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedVector:
    """A length-n vector in the ghost-cell layout of a `ShardedDevBandedMatrix` (`lbm_svec_*`)."""

    def __init__(self, op: "ShardedDevBandedMatrix"):
        self.op = op
        self._v = C.c_void_p()
        op.md.check(_lib.lib().lbm_svec_create(op._s, C.byref(self._v)), "lbm_svec_create")

    def set(self, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=self.op.dtype)
        if x.shape != (self.op.n,):
            raise DimensionMismatch(f"vector has length {x.shape}, operator has n = {self.op.n}")
        self.op.md.check(_lib.lib().lbm_svec_distribute(self._v, x.ctypes.data_as(C.c_void_p)), "lbm_svec_distribute")
        self.op.md.sync()          # x is a pageable host array: do not let it go away under the DMA
        return self

    def fill_uniform(self, seed: int):
        self.op.md.check(_lib.lib().lbm_svec_fill_uniform(self._v, seed & (2**64 - 1)), "lbm_svec_fill_uniform")
        return self

    def get(self) -> np.ndarray:
        y = np.empty(self.op.n, dtype=self.op.dtype)
        self.op.md.check(_lib.lib().lbm_svec_gather(self._v, y.ctypes.data_as(C.c_void_p)), "lbm_svec_gather")
        self.op.md.sync()
        return y

    def __del__(self):
        try:
            if self._v:
                _lib.lib().lbm_svec_destroy(self._v)
        except Exception:
            pass


class ShardedDevBandedMatrix:
    """An n x n BandedMatrix with bandwidths (l,u) row-sharded over the devices of a `MultiDevice` (`lbm_shard_*`)."""

    def __init__(self, md: MultiDevice, n: int, l: int, u: int, dtype=np.float64):
        self.md, self.n, self.l, self.u = md, int(n), int(l), int(u)
        self.dtype = np.dtype(dtype)
        if self.dtype not in _PFX:
            raise LbmArgumentError(f"unsupported element type {dtype}")
        self._s = C.c_void_p()
        md.check(_lib.lib().lbm_shard_create(md.ptr, self.dtype.itemsize, self.n, self.l, self.u, C.byref(self._s)), "lbm_shard_create")

    @classmethod
    def from_host(cls, md, data: np.ndarray, l: int, u: int):
        """`data`: the (l+u+1) x n diagonal-major array (any leading dimension >= l+u+1), Fortran order."""
        data = np.asfortranarray(data)
        op = cls(md, data.shape[1], l, u, data.dtype)
        md.check(_lib.lib().lbm_shard_distribute(op._s, data.ctypes.data_as(C.c_void_p), data.shape[0]), "lbm_shard_distribute")
        md.sync()
        return op

    @classmethod
    def brand(cls, md, n, l, u, dtype=np.float64, seed=1):
        op = cls(md, n, l, u, dtype)
        md.check(_lib.lib().lbm_shard_fill_uniform(op._s, seed & (2**64 - 1)), "lbm_shard_fill_uniform")
        return op

    def bandwidths(self):
        return (self.l, self.u)

    def plan(self, r: int):
        out = (C.c_int64 * 8)()
        p = C.c_void_p()
        self.md.check(_lib.lib().lbm_shard_slab(self._s, r, C.byref(p), out), "lbm_shard_slab")
        return list(out)

    def new_vector(self) -> ShardedVector:
        return ShardedVector(self)

    def mul_(self, y: ShardedVector, x: ShardedVector, alpha=1.0, beta=0.0):
        """`mul!(y, A, x, alpha, beta)` on all devices: ONE C call, one kernel per device."""
        fn = getattr(_lib.lib(), f"lbm_{_PFX[self.dtype]}gbmv_sharded_all")
        self.md.check(fn(self._s, alpha, x._v, beta, y._v), "lbm_gbmv_sharded_all")
        return y

    def __del__(self):
        try:
            if self._s:
                _lib.lib().lbm_shard_destroy(self._s)
        except Exception:
            pass


def chain_mul_(y: ShardedVector, factors, x: ShardedVector, alpha=1.0, beta=0.0):
    """`mul!(y, ApplyMatrix(*, F_0, ..., F_k), x, alpha, beta)` on all devices (`lbm_?gbmv_chain_sharded_all`):
    x in the layout of the LAST factor; the halos of every intermediate are exchanged by the stage that consumes it."""
    md = factors[0].md
    arr = (C.c_void_p * len(factors))(*[f._s.value for f in factors])
    fn = getattr(_lib.lib(), f"lbm_{_PFX[factors[0].dtype]}gbmv_chain_sharded_all")
    md.check(fn(len(factors), arr, alpha, x._v, beta, y._v), "lbm_gbmv_chain_sharded_all")
    return y
