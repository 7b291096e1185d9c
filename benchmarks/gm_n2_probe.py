#!/usr/bin/env python
"""GM at N=2 reads 76 % or 91 % depending on the process: time the per-rank problem of a 2-way column shard on ONE GPU, with
the three slabs allocated back to back (what bench.py does) and with skewed allocations."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L
from lbm_b200.shard import GbmmColumnShard

dev = torch.device("cuda", 0)
nm_, lm = 1 << 26, 8
nbm = 2 * lm + 1


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
    return round(e0.elapsed_time(e1) / reps, 4)


for world in (2, 1, 4):
    for rank in range(min(world, 2)):
        q = GbmmColumnShard.plan(nm_, lm, lm, lm, lm, rank, world)
        for skew in (0, 4096 * 3 + 256, 2 * 1024 * 1024 + 4096 * 5):
            torch.cuda.empty_cache()
            pads = []
            def alloc(rows, cols):
                if skew:
                    pads.append(torch.empty(skew * (len(pads) + 1), dtype=torch.uint8, device=dev))
                return torch.empty((rows, cols), dtype=torch.float64, device=dev)
            A_slab = alloc(q.pb - q.pa, nbm); L.fill_uniform(A_slab, 0x1, offset=nbm * q.pa)
            B_slab = alloc(q.c1 - q.c0, nbm); L.fill_uniform(B_slab, 0x2, offset=nbm * q.c0)
            C_slab = alloc(q.c1 - q.c0, 4 * lm + 1)
            ms = [timeit(lambda: q.local_gbmm(A_slab, B_slab, C_slab)) for _ in range(3)]
            byts = 8 * (q.c1 - q.c0) * (2 * nbm + 4 * lm + 1)
            print(json.dumps({"world": world, "rank": rank, "skew_bytes": skew, "ms": ms, "frac": round(byts / min(ms) / 1e6 / 6551.7, 3),
                              "ptrs_mod_2MB": [int(t.data_ptr() % (1 << 21)) for t in (A_slab, B_slab, C_slab)]}), flush=True)
            del A_slab, B_slab, C_slab, pads
