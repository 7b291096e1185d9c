#!/usr/bin/env python
"""C3 at the size ONE rank of an 8-GPU job holds (1024 x 8192 slab): where do the microseconds go?
(a) host time per call, (b) back-to-back event time, (c) the same applies replayed from a CUDA graph (no host in the loop),
with programmatic dependent launch on / off and the ring-kernel knobs.  JSON lines."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L

PEAK = 6551.7
g, n1 = 8192, int(os.environ.get("N1", "1024"))
D = L.banded_from_diagonals(g, {0: -2.0, 1: 1.0, -1: 1.0})
A1 = L.banded_from_diagonals(n1, {0: -2.0, 1: 1.0, -1: 1.0})
Lap = L.KronSum(A1, D)
x = torch.rand(n1 * g, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
h = L.handle_for(0)
byts = 16 * n1 * g


def events(fn, reps=200, warm=20):
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


def host(fn, reps=2000):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = time.perf_counter() - t
    torch.cuda.synchronize()
    return dt / reps * 1e6


def graph(fn, k=50, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(k):
                fn()
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / k * 1e3


def out(name, us):
    print(json.dumps({"config": name, "us": round(us, 2), "frac": round(byts / us / 1e3 / PEAK, 3), "ideal_us": round(byts / PEAK / 1e3, 2)}), flush=True)


fn = lambda: Lap.mul_(y, x)
small = L.KronSum(L.banded_from_diagonals(64, {0: -2.0, 1: 1.0, -1: 1.0}), L.banded_from_diagonals(64, {0: -2.0, 1: 1.0, -1: 1.0}))
xs = torch.rand(64 * 64, dtype=torch.float64, device="cuda")
ys = torch.empty_like(xs)
print(json.dumps({"host_us_per_call_KronSum.mul_": round(host(lambda: small.mul_(ys, xs)), 2)}), flush=True)
for pdl in (0, 1):
    h.set_tuning(pdl=pdl)
    tag = "pdl on" if pdl == 1 else "pdl off"
    out(f"events back-to-back {tag}", events(fn))
    try:
        out(f"graph replay {tag}", graph(fn))
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"graph": "failed", "why": str(e)[:200]}), flush=True)
        break
for tr in (0, 512):
    for tc in (4, 8, 16):
        for st in (2, 3):
            for cps in (0, 3, 4, 5, 6):
                h.set_tuning(kron_tr=tr, kron_tc=tc, kron_stages=st, kron_ctas_per_sm=cps, pdl=0)
                try:
                    out(f"graph replay tr={tr} tc={tc} stages={st} cps={cps}", graph(fn, 30, 10))
                except Exception as e:  # noqa: BLE001
                    print(json.dumps({"graph": "failed", "cfg": [tr, tc, st, cps], "why": str(e)[:200]}), flush=True)
h.set_tuning(kron_tr=0, kron_tc=0, kron_stages=0, kron_ctas_per_sm=0)
