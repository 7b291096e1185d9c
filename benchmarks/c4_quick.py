#!/usr/bin/env python
"""C4 (l=u=128, n=2^20) with a list of (gbmm_variant, gbmm_tile) tuning pairs: python benchmarks/c4_quick.py 0:5 0:205 0:305
Each line carries the SM clock / power sampled by NVML DURING the timed launches (FP64 tensor work sits near the power cap)."""
import os, sys, json, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L

log2n = int(os.environ.get("C4_LOG2N", "20"))
n = 1 << log2n
A = L.brand(n, n, 128, 128, seed=1)
B = L.brand(n, n, 128, 128, seed=2)
Cm = L.BandedMatrix(torch.empty((n, 513), dtype=torch.float64, device="cuda"), n, 256, 256)
h = L.handle_for(0)
flops = 2 * n * 257 * 257
pairs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(0, 0)]


class Sampler:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv = pynvml
        self.hd = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.sm, self.pw, self.reasons = [], [], 0
        self.stop = threading.Event()

    def run(self):
        while not self.stop.is_set():
            self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.hd, self.nv.NVML_CLOCK_SM))
            self.pw.append(self.nv.nvmlDeviceGetPowerUsage(self.hd) / 1000.0)
            self.reasons |= self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.hd)
            time.sleep(0.002)


for rep in range(2):
    for v, t in pairs:
        h.set_tuning(gbmm_variant=v, gbmm_tile=t)
        for _ in range(2):
            L.gbmm(A, B, out=Cm)
        torch.cuda.synchronize()
        s = Sampler()
        th = threading.Thread(target=s.run, daemon=True)
        th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 20
        for _ in range(reps):
            L.gbmm(A, B, out=Cm)
        e1.record()
        torch.cuda.synchronize()
        s.stop.set()
        th.join()
        ms = e0.elapsed_time(e1) / reps
        sm = sorted(s.sm)
        print(json.dumps({"variant": v, "tile": t, "ms": round(ms, 3), "TFLOPs": round(flops / ms / 1e9, 2),
                          "sm_mhz_median": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
                          "power_w_max": round(max(s.pw), 1) if s.pw else None, "throttle_reasons": hex(s.reasons)}), flush=True)
