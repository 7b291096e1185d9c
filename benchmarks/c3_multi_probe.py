#!/usr/bin/env python
"""C3 at N ranks with 1024 grid columns per rank (what one rank of the 8-GPU run holds): where does the time of an apply go?
torchrun ... benchmarks/c3_multi_probe.py   -- JSON lines from rank 0 (per-rank ms listed)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import lbm_b200 as L
from lbm_b200.shard import RowShard, ShardedKronSum, init_comm

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
init_comm(local)
h = L.handle_for(local)
g = 8192
cols = int(os.environ.get("C3_COLS_PER_RANK", "1024")) * world
D2 = L.banded_from_diagonals(g, {0: -2.0, 1: 1.0, -1: 1.0}, device=dev)
D1 = L.banded_from_diagonals(cols, {0: -2.0, 1: 1.0, -1: 1.0}, device=dev)
p = RowShard.plan(cols, 1, 1, rank, world)
sh = ShardedKronSum(cols, g, 1, 1, D1.data[p.c0:p.c1].contiguous(), D2, rank, world)
xb = sh.new_buffer(torch.float64, dev)
L.fill_uniform(sh.owned(xb).view(-1), 5, offset=p.r0 * g)
y = torch.empty((p.m_loc, g), dtype=torch.float64, device=dev)
# the same slab as a purely local operator (no neighbours, plain kernel)
A_loc = L.banded_from_diagonals(p.m_loc, {0: -2.0, 1: 1.0, -1: 1.0}, device=dev)
loc = L.KronSum(A_loc, D2)
xl = torch.rand(p.m_loc * g, dtype=torch.float64, device=dev)
yl = torch.empty_like(xl)
byts = 16 * g * cols


def run(name, fn, reps, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host_us = (time.perf_counter() - t0) / reps * 1e6
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3, host_us], dtype=torch.float64, device=dev)
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        us = [round(float(v[0]), 2) for v in allt]
        print(json.dumps({"config": name, "reps": reps, "us_per_rank": us, "us_max": max(us), "host_enqueue_us_per_rank": [round(float(v[1]), 2) for v in allt],
                          "frac_of_aggregate": round(byts / max(us) / 1e3 / (6551.7 * world), 3)}), flush=True)


for reps in (20, 200):
    h.set_tuning(exchange_mode=0, exchange_fuse=0)
    run("sharded, fused exchange", lambda: sh.mul_(y, xb), reps)
    h.set_tuning(exchange_mode=0, exchange_fuse=2)
    run("sharded, two-kernel exchange", lambda: sh.mul_(y, xb), reps)
    h.set_tuning(exchange_mode=0, exchange_fuse=0)
    run("local operator of the same slab, no neighbours (KronSum.mul_)", lambda: loc.mul_(yl, xl), reps)
cfgs = [tuple(int(v) for v in c.split(":")) for c in os.environ.get("C3_CFGS", "4:2:3,8:1:3,8:2:2,8:1:4,16:1:2").split(",")]
for tc, st, cps in cfgs:
    h.set_tuning(kron_tc=tc, kron_stages=st, kron_ctas_per_sm=cps)
    run(f"sharded, fused exchange, tc={tc} stages={st} cps={cps}", lambda: sh.mul_(y, xb), 200)
h.set_tuning(kron_tc=0, kron_stages=0, kron_ctas_per_sm=0)
dist.barrier()
dist.destroy_process_group()
