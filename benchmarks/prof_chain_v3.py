import os, sys
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import lbm_b200 as L
h = L.handle_for(0)
batch, n = 4096, 4096
fs = [torch.empty((batch, n, 9), dtype=torch.float32, device="cuda") for _ in range(3)]
for i, f in enumerate(fs):
    L.fill_uniform(f, 1 + i)
h.set_tuning(chain_variant=3)
for _ in range(2):
    D = L.chain3_batched(*fs, [4, 4, 4], [4, 4, 4])
torch.cuda.synchronize()
print("ok", float(D[0, 0, 12]))
