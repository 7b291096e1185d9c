#!/usr/bin/env python
"""banded x dense with a WIDE band (l=u=128), 8 right-hand sides: one pass over A (gbmv_wide_n_multi_kernel) vs column by column."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L

PEAK = 6551.7
n, nrhs, l = 1 << 22, 8, 128
A = L.brand(n, n, l, l, seed=1)
X = torch.empty((nrhs, n), dtype=torch.float64, device="cuda"); L.fill_uniform(X, 9)
Y = torch.empty_like(X)
byts = 8 * n * ((2 * l + 1) + 2 * nrhs)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


h = L.handle_for(0)
for chunk, stages in ((16, 3), (16, 4), (32, 2), (48, 2)):
    h.set_tuning(gbmvw_chunk=chunk, gbmvw_stages=stages)
    ms = timeit(lambda: L.mul_(Y.t(), A, X.t()))
    print(json.dumps({"config": f"MM wide multi chunk={chunk} stages={stages}", "ms": round(ms, 3), "frac_hbm_measured": round(byts / ms / 1e6 / PEAK, 3)}), flush=True)
h.set_tuning(gbmvw_chunk=0, gbmvw_stages=0)
for name, fn in (("lbm_dgbmv_multi (one pass over A)", lambda: L.mul_(Y.t(), A, X.t())),
                 ("column by column (8 gbmv launches)", lambda: [L.mul_(Y[c], A, X[c]) for c in range(nrhs)])):
    ms = timeit(fn)
    print(json.dumps({"config": f"MM banded*dense f64 n=2^22 l=u=128 nrhs=8, {name}", "ms": round(ms, 3), "GBs": round(byts / ms / 1e6, 1),
                      "frac_hbm_measured": round(byts / ms / 1e6 / PEAK, 3)}), flush=True)
# op T (A' * X): gbmv_wide_t_multi_kernel (round 2) vs column by column
for chunk, stages, cps in ((0, 0, 0), (32, 2, 0), (32, 3, 1), (64, 2, 0), (64, 3, 0), (32, 4, 0)):
    h.set_tuning(gbmvw_chunk=chunk, gbmvw_stages=stages, gbmvw_ctas_per_sm=cps)
    ms = timeit(lambda: L.mul_(Y.t(), A.T, X.t()))
    print(json.dumps({"config": f"MM wide multi op T chunk={chunk} stages={stages} ctas={cps}", "ms": round(ms, 3), "frac_hbm_measured": round(byts / ms / 1e6 / PEAK, 3)}), flush=True)
h.set_tuning(gbmvw_chunk=0, gbmvw_stages=0, gbmvw_ctas_per_sm=0)
ms = timeit(lambda: [L.mul_(Y[c], A.T, X[c]) for c in range(nrhs)])
print(json.dumps({"config": "MM op T column by column (8 gbmv launches)", "ms": round(ms, 3), "frac_hbm_measured": round(byts / ms / 1e6 / PEAK, 3)}), flush=True)
