#!/usr/bin/env python
"""A/B of ring-kernel configurations for C3 slabs, alternating, 3 rounds: N1=... python benchmarks/kron_ab.py tr:tc:stages:cps ..."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L

PEAK = 6551.7
g = 8192
h = L.handle_for(0)
D = L.banded_from_diagonals(g, {0: -2.0, 1: 1.0, -1: 1.0})
cfgs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]]
for n1 in [int(v) for v in os.environ.get("N1", "1024,8192").split(",")]:
    A1 = L.banded_from_diagonals(n1, {0: -2.0, 1: 1.0, -1: 1.0})
    Lap = L.KronSum(A1, D)
    x = torch.rand(n1 * g, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    byts = 16 * n1 * g
    s = torch.cuda.Stream()
    res = {c: [] for c in cfgs}
    with torch.cuda.stream(s):
        for rnd in range(4):
            for c in cfgs:
                h.set_tuning(kron_tr=c[0], kron_tc=c[1], kron_stages=c[2], kron_ctas_per_sm=c[3])
                Lap.mul_(y, x)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=s):
                    for _ in range(20):
                        Lap.mul_(y, x)
                gr.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    gr.replay()
                e1.record()
                torch.cuda.synchronize()
                res[c].append(e0.elapsed_time(e1) / 200 * 1e3)
    for c in cfgs:
        v = sorted(res[c])
        print(json.dumps({"n1": n1, "cfg": "tr:tc:stages:cps=%d:%d:%d:%d" % c, "us_min": round(v[0], 2), "us_median": round((v[1] + v[2]) / 2, 2),
                          "us_max": round(v[-1], 2), "frac_median": round(byts / ((v[1] + v[2]) / 2) / 1e3 / PEAK, 3)}), flush=True)
    del A1, Lap, x, y
h.set_tuning(kron_tr=0, kron_tc=0, kron_stages=0, kron_ctas_per_sm=0)
