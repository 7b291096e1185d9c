import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/lazybandedmatrices.jl_b200")
import torch, lbm_b200 as L
n = 1 << 20
A = L.brand(n, n, 128, 128, seed=1); B = L.brand(n, n, 128, 128, seed=2)
Cm = L.BandedMatrix(torch.empty((n, 513), dtype=torch.float64, device="cuda"), n, 256, 256)
h = L.handle_for(0)
def t(reps=3):
    L.gbmm(A, B, out=Cm); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): L.gbmm(A, B, out=Cm)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for v, tile in ((22, 1), (22, 101), (20, 5), (20, 105)):   # +100: every cp.async issued twice (2x L2->smem traffic)
    h.set_tuning(gbmm_variant=v, gbmm_tile=tile)
    print(v, tile, round(t(), 3), "ms")
