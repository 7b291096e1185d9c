#!/usr/bin/env python
"""Tuning sweep on one GPU: GB/s of the streaming kernels over their launch knobs.
    python benchmarks/sweep.py [--log2n 26] > gpurun_out/sweep.json"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import lbm_b200 as L  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=26)
    ap.add_argument("--what", default="gbmv1,gbmv8,tri,wide,copy")
    a = ap.parse_args()
    n = 1 << a.log2n
    h = L.handle_for(0)
    out = []
    what = a.what.split(",")
    if "copy" in what:
        src = torch.empty(n * 4, dtype=torch.float64, device="cuda")
        dst = torch.empty_like(src)
        ms = timeit(lambda: dst.copy_(src))
        out.append({"k": "torch_copy", "gbs": 2 * src.numel() * 8 / ms / 1e6})
        del src, dst
    for name, (l, u) in (("gbmv1", (1, 1)), ("gbmv8", (8, 8))):
        if name not in what:
            continue
        A = L.brand(n, n, l, u)
        x = torch.empty(n, dtype=torch.float64, device="cuda")
        L.fill_uniform(x, 9)
        y = torch.empty_like(x)
        byts = 8 * n * (l + u + 3)
        tiles = (256, 512, 1024, 2048) if l == 1 else (256, 512)
        for tile, st, ct in itertools.product(tiles, (1, 2, 3, 4), (1, 2, 3, 4, 6)):
            h.set_tuning(gbmv_tile=tile, gbmv_stages=st, gbmv_ctas_per_sm=ct)
            ms = timeit(lambda: L.mul_(y, A, x), reps=5, warm=2)
            out.append({"k": name, "tile": tile, "stages": st, "ctas": ct, "ms": ms, "gbs": byts / ms / 1e6})
        h.set_tuning(gbmv_tile=0, gbmv_stages=0, gbmv_ctas_per_sm=0, gbmv_force=1)
        ms = timeit(lambda: L.mul_(y, A, x), reps=5, warm=2)
        out.append({"k": name, "direct": True, "ms": ms, "gbs": byts / ms / 1e6})
        h.set_tuning(gbmv_force=0)
        ms = timeit(lambda: L.mul_(y, A.T, x), reps=5, warm=2)
        out.append({"k": name, "trans": True, "ms": ms, "gbs": byts / ms / 1e6})
        del A, x, y
        torch.cuda.empty_cache()
    if "tri" in what:
        d = torch.empty(n, dtype=torch.float64, device="cuda")
        e = torch.empty(n - 1, dtype=torch.float64, device="cuda")
        e2 = torch.empty(n - 1, dtype=torch.float64, device="cuda")
        x = torch.empty(n, dtype=torch.float64, device="cuda")
        for t, s in ((d, 1), (e, 2), (e2, 3), (x, 4)):
            L.fill_uniform(t, s)
        y = torch.empty_like(x)
        for nm, A, arrays in (("symtri", L.SymTridiagonal(d, e), 4), ("tri", L.Tridiagonal(e, d, e2), 5),
                              ("bidiagU", L.Bidiagonal(d, e, "U"), 4), ("bidiagL", L.Bidiagonal(d, e, "L"), 4)):
            for ct in (2, 4, 8, 16, 32):
                h.set_tuning(tri_ctas_per_sm=ct)
                ms = timeit(lambda: L.mul_(y, A, x), reps=5, warm=2)
                out.append({"k": nm, "ctas": ct, "ms": ms, "gbs": arrays * 8 * n / ms / 1e6})
        h.set_tuning(tri_ctas_per_sm=0)
        del d, e, e2, x, y
        torch.cuda.empty_cache()
    if "wide" in what:
        nw = 1 << min(a.log2n, 23)
        l = u = 128
        A = L.brand(nw, nw, l, u)
        x = torch.empty(nw, dtype=torch.float64, device="cuda")
        L.fill_uniform(x, 9)
        y = torch.empty_like(x)
        byts = 8 * nw * (l + u + 3)
        for ch, st in itertools.product((32, 64), (2, 3, 4, 6)):
            h.set_tuning(gbmvw_chunk=ch, gbmvw_stages=st)
            try:
                ms = timeit(lambda: L.mul_(y, A, x), reps=3, warm=1)
                out.append({"k": "wide128", "chunk": ch, "stages": st, "ms": ms, "gbs": byts / ms / 1e6})
            except Exception as ex:
                out.append({"k": "wide128", "chunk": ch, "stages": st, "error": str(ex)[:100]})
        h.set_tuning(gbmvw_chunk=0, gbmvw_stages=0)
        ms = timeit(lambda: L.mul_(y, A.T, x), reps=3, warm=1)
        out.append({"k": "wide128", "trans": True, "ms": ms, "gbs": byts / ms / 1e6})
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
