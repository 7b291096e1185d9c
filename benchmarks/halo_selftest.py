#!/usr/bin/env python
"""ONE GPU: the HALO (fused-exchange) instantiation of the streaming kernels with NO neighbours (exchange_fuse = 3) against
the plain instantiation -- what the variant itself costs (registers, rotated tile order, per-tile overlap tests)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import lbm_b200 as L  # noqa: E402
from lbm_b200.shard import RowShard, ShardedBanded, ShardedKronSum  # noqa: E402


def timeit(fn, reps=100, warm=10):
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


h = L.handle_for(0)
dev = torch.device("cuda", 0)
g = 8192
D = L.banded_from_diagonals(g, {0: -2.0, 1: 1.0, -1: 1.0})
for n1 in (1024, 4096, 8192):
    A = L.banded_from_diagonals(n1, {0: -2.0, 1: 1.0, -1: 1.0})
    p = RowShard.plan(n1, 1, 1, 0, 1)
    sh = ShardedKronSum(n1, g, 1, 1, A.data[p.c0:p.c1].contiguous(), D, 0, 1)
    xb = sh.new_buffer(torch.float64, dev)
    xb.uniform_()
    y = torch.empty((p.m_loc, g), dtype=torch.float64, device=dev)
    res = {}
    for fuse, sub in ((0, 0), (3, 0), (3, 2), (3, 4), (3, 6), (0, 0), (3, 0)):
        h.set_tuning(exchange_fuse=fuse, halo_timeout_ms=sub)
        res.setdefault((fuse, sub), []).append(round(timeit(lambda: sh.mul_(y, xb)), 2))
    print(json.dumps({"kronsum n1": n1, "plain_us": res[(0, 0)], "halo_variant_us": res[(3, 0)], "no_rotation": res[(3, 2)], "no_patch_test": res[(3, 4)],
                      "neither": res[(3, 6)]}), flush=True)
for (n, l) in ((1 << 24, 1), (1 << 24, 8), (1 << 26, 1)):
    sh = ShardedBanded.brand(n, l, l, 0, 1, device=dev)
    xb = sh.plan.new_vector(torch.float64, dev)
    xb.uniform_()
    y = torch.empty(n, dtype=torch.float64, device=dev)
    fn = getattr(L._lib.lib(), "lbm_dgbmv_sharded")
    import ctypes as C
    args = (h.ptr, n, l, l, C.c_double(1.0), C.c_void_p(sh.data.data_ptr()), 2 * l + 1, C.c_void_p(xb.data_ptr()), C.c_double(0.0),
            C.c_void_p(y.data_ptr()), 1)
    res = {}
    for fuse, sub in ((0, 0), (3, 0), (3, 2), (3, 4), (3, 6), (0, 0), (3, 0)):
        h.set_tuning(exchange_fuse=fuse, halo_timeout_ms=sub)
        h.use_torch_stream()
        res.setdefault((fuse, sub), []).append(round(timeit(lambda: fn(*args)), 2))
    print(json.dumps({"gbmv n": n, "l": l, "plain_us": res[(0, 0)], "halo_variant_us": res[(3, 0)], "no_rotation": res[(3, 2)],
                      "no_patch_test": res[(3, 4)], "neither": res[(3, 6)]}), flush=True)
h.set_tuning(exchange_fuse=0, halo_timeout_ms=0)
