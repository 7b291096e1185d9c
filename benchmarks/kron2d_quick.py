import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch, lbm_b200 as L
g = 8192
D = L.banded_from_diagonals(g, {0: -2.0, 1: 1.0, -1: 1.0})
K = L.BlockKron(D, D)
x = torch.empty(g * g, dtype=torch.float64, device="cuda"); L.fill_uniform(x, 5)
y = torch.empty_like(x)
h = L.handle_for(0)
for force in (0, 1):
    h.set_tuning(kron_force_tile=force)
    for _ in range(3): K.mul_(y, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): K.mul_(y, x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"config": f"BlockKron(D,D)*x 8192^2 (kron_force_tile=0: one-pass ring kernel, 1: two-pass tile kernel), kron_force_tile={force}", "ms": ms, "GBs_alg_2N": 16 * g * g / ms / 1e6, "GBs_two_pass_4N": 32 * g * g / ms / 1e6}))
