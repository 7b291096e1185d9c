#!/usr/bin/env python
"""(1,1) vs (0,2) vs (2,0) rectangular slabs (rank 0 vs the other ranks of a row shard), interleaved repeats."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L

h = L.handle_for(0)
n = 1 << int(os.environ.get("LOG2N", "24"))


def timeit(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


ops = {}
for (kl, ku) in ((1, 1), (0, 2), (2, 0)):
    data = torch.rand((n + 2, 3), dtype=torch.float64, device="cuda")
    ops[(kl, ku)] = (L.BandedMatrix(data, n, kl, ku), torch.rand(n + 2, dtype=torch.float64, device="cuda"),
                     torch.empty(n, dtype=torch.float64, device="cuda"))
configs = [("default", {}), ("cps2", dict(gbmv_ctas_per_sm=2)), ("cps4", dict(gbmv_ctas_per_sm=4)), ("stages3", dict(gbmv_stages=3)),
           ("tile2048", dict(gbmv_tile=2048)), ("tile768 st3", dict(gbmv_tile=768, gbmv_stages=3)),
           ("tile1280 cps4", dict(gbmv_tile=1280, gbmv_ctas_per_sm=4)), ("delay300", dict(gbmv_skew=-300)), ("delay1000", dict(gbmv_skew=-1000)),
           ("tile512 cps6", dict(gbmv_tile=512, gbmv_ctas_per_sm=6))]
res = {}
for rep in range(3):
    for name, kw in configs:
        h.set_tuning(gbmv_tile=0, gbmv_stages=0, gbmv_ctas_per_sm=0, gbmv_skew=0)
        h.set_tuning(**kw)
        for key, (A, x, y) in ops.items():
            res.setdefault((name, key), []).append(round(timeit(lambda: L.mul_(y, A, x)), 1))
h.set_tuning(gbmv_tile=0, gbmv_stages=0, gbmv_ctas_per_sm=0, gbmv_skew=0)
for name, _ in configs:
    print(json.dumps({"config": name, **{str(k): res[(name, k)] for k in ops}}), flush=True)
