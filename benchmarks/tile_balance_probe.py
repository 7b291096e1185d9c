import os, sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/lazybandedmatrices.jl_b200")
import torch, lbm_b200 as L
h = L.handle_for(0)
def timeit(fn, reps=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps * 1e3, 2)
for log2n in (24, 25):
    n = 1 << log2n
    for (kl, ku) in ((1, 1), (0, 2), (2, 0)):
        A = L.BandedMatrix(torch.rand((n + 2, 3), dtype=torch.float64, device="cuda"), n, kl, ku)
        x = torch.rand(n + 2, dtype=torch.float64, device="cuda"); y = torch.empty(n, dtype=torch.float64, device="cuda")
        r = {}
        for rep in range(2):
            for name, kw in (("default", {}), ("tile1280 cps4", dict(gbmv_tile=1280, gbmv_ctas_per_sm=4)), ("tile1024 cps4", dict(gbmv_tile=1024, gbmv_ctas_per_sm=4))):
                h.set_tuning(gbmv_tile=0, gbmv_stages=0, gbmv_ctas_per_sm=0); h.set_tuning(**kw)
                r.setdefault(name, []).append(timeit(lambda: L.mul_(y, A, x)))
        h.set_tuning(gbmv_tile=0, gbmv_stages=0, gbmv_ctas_per_sm=0)
        print(json.dumps({"n": f"2^{log2n}", "bands": [kl, ku], **r}), flush=True)
