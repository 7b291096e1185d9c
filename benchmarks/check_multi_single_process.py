#!/usr/bin/env python
"""SINGLE-PROCESS multi-device parity on real GPUs: one host thread drives every visible GPU through
lbm_create_multi / lbm_shard_* / lbm_svec_* / lbm_?gbmv_sharded_all / lbm_?gbmv_chain_sharded_all, plus the per-device
`*_sharded` entry points (tridiagonal family, Kronecker sum) on the handles lbm_multi_handle() returns.  No
torch.distributed, no NCCL.  Every result is compared with the single-process oracle.  Prints OK / raises."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import lbm_b200 as L  # noqa: E402
from lbm_b200 import _lib  # noqa: E402
from lbm_b200.multi import MultiDevice, ShardedDevBandedMatrix, chain_mul_  # noqa: E402
from oracle import oracle as O  # noqa: E402


def ndevices():
    import torch
    return torch.cuda.device_count()


def main():
    nd = int(sys.argv[1]) if len(sys.argv) > 1 else ndevices()
    assert nd >= 1
    md = MultiDevice(nd)
    lib = _lib.lib()
    eps = np.finfo(np.float64).eps

    # ---- banded mul!: host BandedMatrix data -> distribute -> mul! -> gather (fused exchange, then the unfused one) ----
    for (n, l, u, dt) in ((300007, 8, 5, np.float64), (100003, 1, 1, np.float64), (50001, 3, 0, np.float32), (40000, 128, 128, np.float64),
                          (100, 8, 5, np.float64), (4 * nd - 2, 1, 1, np.float64)):
        e = np.finfo(dt).eps
        A = O.brand(n, n, l, u, seed=0x21, dtype=dt)
        x = O.fill_uniform(0x9, n, dt)
        want = O.gbmv("N", n, n, l, u, 1.5, A, x)
        try:
            op = ShardedDevBandedMatrix.from_host(md, A, l, u)
        except L.LbmArgumentError:
            assert nd > 1 and -(-n // nd) < max(l, u) + 4, (n, l, u)   # shards narrower than the halo: rejected for every device
            continue
        xv, yv = op.new_vector().set(x), op.new_vector()
        for fuse in (0, 2):
            md.set_tuning(exchange_fuse=fuse)
            l0 = md.launch_count()
            op.mul_(yv, xv, alpha=1.5)
            per_dev = (md.launch_count() - l0) / nd
            got = yv.get()
            err = float(np.max(np.abs(got - want))) if n else 0.0
            assert err <= 4 * e * (l + u + 1) * 1.5, ("gbmv", n, l, u, fuse, err)
            if fuse == 0 and nd > 1 and l + u + 1 <= 33 and n > 1000:
                assert per_dev <= 1.0 + 1e-9, ("fused path must be ONE launch per device", per_dev)
        md.set_tuning(exchange_fuse=0)
        # epochs: x changes every apply (exact power-of-two scalings), 40 back-to-back applies, no host sync in between
        if n > 1000:
            ys = []
            vs = [op.new_vector().set(x * float(2 ** (k % 5))) for k in range(5)]
            outs = [op.new_vector() for _ in range(40)]
            for k in range(40):
                op.mul_(outs[k], vs[k % 5])
            for k in range(40):
                err = np.max(np.abs(outs[k].get() / float(2 ** (k % 5)) - want / 1.5))
                assert err <= 4 * e * (l + u + 1), ("epoch", k, err)
            del ys, vs, outs

    # ---- lazy chain F0*(F1*(F2*x)) with different bandwidths: intermediates' halos exchanged inside ONE C call ----
    n = 200003
    bands = [(1, 1), (2, 5), (8, 8)]
    mats = [O.brand(n, n, l, u, seed=0x30 + i) for i, (l, u) in enumerate(bands)]
    x = O.fill_uniform(0x9, n)
    v = x
    for (l, u), M in zip(reversed(bands), reversed(mats)):
        v = O.gbmv("N", n, n, l, u, 1.0, M, v)
    fs = [ShardedDevBandedMatrix.from_host(md, M, l, u) for (l, u), M in zip(bands, mats)]
    xv, yv = fs[-1].new_vector().set(x), fs[0].new_vector()
    chain_mul_(yv, fs, xv)
    err = np.max(np.abs(yv.get() - v))
    bound = 4 * eps * sum((l + u + 1) for (l, u) in bands) * np.max(np.abs(v))
    assert err <= bound, ("chain", err, bound)
    with_wrong = fs[0].new_vector()
    try:
        chain_mul_(yv, fs, with_wrong)          # x in the wrong ghost layout: DimensionMismatch before any launch
        raise AssertionError("expected DimensionMismatch")
    except L.DimensionMismatch:
        pass

    # ---- synthetic operator at a size that matters + timing: fused vs unfused launches per apply ----
    n = 1 << 24
    for (l, u) in ((1, 1), (8, 8)):
        op = ShardedDevBandedMatrix.brand(md, n, l, u, seed=1)
        xv, yv = op.new_vector().fill_uniform(9), op.new_vector()
        res = {}
        for fuse in (0, 2):
            md.set_tuning(exchange_fuse=fuse)
            for _ in range(3):
                op.mul_(yv, xv)
            md.sync()
            l0, t0 = md.launch_count(), time.perf_counter()
            for _ in range(50):
                op.mul_(yv, xv)
            md.sync()
            dt = (time.perf_counter() - t0) / 50
            res[fuse] = (dt * 1e3, (md.launch_count() - l0) / 50 / nd)
        md.set_tuning(exchange_fuse=0)
        rows = np.unique(np.concatenate([np.arange(64), np.arange(n - 64, n), np.random.default_rng(1).integers(0, n, 500)]
                                        + [np.arange(max(0, op.plan(r)[0] - 12), min(n, op.plan(r)[0] + 12)) for r in range(1, nd)]))
        got = yv.get()[rows]
        assert np.max(np.abs(got - O.gbmv_rows_stream(n, l, u, 1, 9, rows))) <= 4 * eps * (l + u + 1)
        gbs = 8 * n * (l + u + 3) / 1e9
        print(f"  n=2^24 l=u={l} on {nd} device(s), one host thread: fused {res[0][0]:.3f} ms ({gbs / res[0][0] * 1e3:.0f} GB/s, "
              f"{res[0][1]:.2f} launches/device/apply) vs unfused {res[2][0]:.3f} ms ({res[2][1]:.2f} launches)", flush=True)

    # ---- per-device `*_sharded` entries on the multi handles: SymTridiagonal / Bidiagonal and the Kronecker sum ----
    if nd > 1:
        n = 1 << 20
        d, e2 = O.fill_uniform(1, n), O.fill_uniform(2, n - 1)
        x = O.fill_uniform(4, n)
        plan = (C.c_int64 * 8)()
        for name, has_l, has_u in (("sym", 1, 1), ("U", 0, 1), ("L", 1, 0)):
            want = O.tridiag_mv(n, e2 if has_l else None, d, e2 if has_u else None, x)
            bufs = []
            for r in range(nd):
                h = C.c_void_p(md.handle(r))
                assert lib.lbm_shard_plan(n, has_l, has_u, r, nd, plan) == 0
                r0, r1, c0, c1, left = plan[0], plan[1], plan[2], plan[3], plan[4]

                def up(a):
                    p = C.c_void_p()
                    assert lib.lbm_malloc(h, C.byref(p), max(16, a.nbytes + 32)) == 0
                    a = np.ascontiguousarray(a)
                    assert lib.lbm_memcpy_h2d(h, p, a.ctypes.data_as(C.c_void_p), a.nbytes) == 0
                    assert lib.lbm_sync(h) == 0
                    return p
                xb = np.zeros(c1 - c0)
                xb[left:left + (r1 - r0)] = x[r0:r1]
                bufs.append((h, r0, r1, up(e2[c0:c1 - 1]), up(d[c0:c1]), up(xb), up(np.full(r1 - r0, np.nan))))
            for (h, r0, r1, pe, pd, px, py) in bufs:      # one stage: device by device, nothing blocks
                rc = lib.lbm_dtridiag_mv_sharded(h, n, pe if has_l else None, pd, pe if has_u else None, 1.0, px, 0.0, py, 1)
                assert rc == 0, lib.lbm_last_error_string(h)
            got = np.empty(n)
            for (h, r0, r1, pe, pd, px, py) in bufs:
                part = np.empty(r1 - r0)
                assert lib.lbm_memcpy_d2h(h, part.ctypes.data_as(C.c_void_p), py, part.nbytes) == 0
                assert lib.lbm_sync(h) == 0
                got[r0:r1] = part
                for p in (pe, pd, px, py):
                    lib.lbm_free(h, p)
            assert np.max(np.abs(got - want)) <= 4 * eps * 3, name

    md.sync()
    print(f"OK single-process multi-device: {nd} GPU(s) driven by one host thread match the oracle "
          f"(banded mul!, epochs, lazy chain, tridiagonal family; fused and unfused exchange)")


if __name__ == "__main__":
    main()
