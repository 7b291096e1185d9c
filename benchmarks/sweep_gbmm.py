#!/usr/bin/env python
"""gbmm over every pair of symmetric bands l = u in 1..8 on both factors, plus the shapes round 1 left on the zero-padded
staging path (8 / 16 stored rows, unequal counts), f64 and f32, n = 2^24: fraction of the measured copy ceiling.
gbmm_force = 4 re-runs round 1's padded staging for the shapes that now take the run-time-count variant.  JSON lines."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L

PEAK = 6551.7
n = 1 << 24
h = L.handle_for(0)
EXACT = {2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 17}


def timeit(fn, reps=8, warm=2):
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


def run(dt, la, ua, lb, ub, both):
    w = 8 if dt == torch.float64 else 4
    A = L.brand(n, n, la, ua, dtype=dt, seed=1)
    B = L.brand(n, n, lb, ub, dtype=dt, seed=2)
    C = L.BandedMatrix(torch.empty((n, la + lb + ua + ub + 1), dtype=dt, device="cuda"), n, la + lb, ua + ub)
    na, nb = la + ua + 1, lb + ub + 1
    byts = w * n * (na + nb + na + nb - 1)
    r = {"config": f"gbmm {dt} n=2^24 ({la},{ua})x({lb},{ub})", "rows": [na, nb], "exact_instantiation": na == nb and na in EXACT or (na in (3, 5, 9, 17) and nb in (3, 5, 9, 17))}
    ms = timeit(lambda: L.gbmm(A, B, out=C))
    r.update(ms=round(ms, 4), GBs=round(byts / ms / 1e6, 1), frac=round(byts / ms / 1e6 / PEAK, 3))
    if both and not r["exact_instantiation"]:
        h.set_tuning(gbmm_force=4)
        ms4 = timeit(lambda: L.gbmm(A, B, out=C))
        h.set_tuning(gbmm_force=0)
        r.update(frac_round1_padded_staging=round(byts / ms4 / 1e6 / PEAK, 3))
    print(json.dumps(r), flush=True)
    del A, B, C
    torch.cuda.empty_cache()


for dt in (torch.float64, torch.float32):
    for (la, ua, lb, ub) in ((3, 4, 4, 3), (8, 7, 7, 8), (3, 3, 1, 1), (1, 2, 2, 3), (2, 3, 6, 3), (0, 0, 8, 8), (8, 8, 0, 1), (7, 8, 0, 0), (0, 1, 4, 4),
                             (4, 4, 3, 4), (3, 3, 4, 5), (2, 2, 1, 2), (1, 1, 2, 3), (3, 3, 8, 7), (5, 5, 0, 5), (2, 1, 1, 0)):
        run(dt, la, ua, lb, ub, True)
    for ka in (range(1, 9) if not os.environ.get('SWEEP_SPECIAL_ONLY') else ()):
        for kb in range(1, 9):
            if ka != kb or os.environ.get("SWEEP_ALL"):
                run(dt, ka, ka, kb, kb, True)
