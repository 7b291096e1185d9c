#!/usr/bin/env python
"""BASELINE configs C2-C5 (+ M-128) across N GPUs (run under torchrun, one rank per GPU): aggregate
algorithmic GB/s or TFLOP/s, max-over-ranks device time, exchange inside the timed region."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import lbm_b200 as L  # noqa: E402
from lbm_b200.shard import (GbmmColumnShard, RowShard, ShardedBanded, ShardedKronSum, ShardedTridiagonal,  # noqa: E402
                            init_comm, split_batch)

PEAK = 6551.7
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def main():
    only = None
    if "--only" in sys.argv:
        only = sys.argv[sys.argv.index("--only") + 1].split(",")
    want = lambda k: only is None or k in only  # noqa: E731
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        init_comm(local)
    h = L.handle_for(local)
    peer = world > 1 and bool(L._lib.lib().lbm_peer_active(h.ptr))
    # ghost transports to report for the configs that exchange: peer-memory mailboxes (default) and NCCL send/recv
    modes = (((0, 0, "peer mailboxes, exchange fused into the launch"), (0, 2, "peer mailboxes, two-kernel exchange"), (1, 0, "ncclSend/Recv"))
             if peer else ((1, 0, "ncclSend/Recv" if world > 1 else "1 GPU"),))

    def both(name, fn, byts, reps=10, warm=3):
        for mode, fuse, label in modes:
            h.set_tuning(exchange_mode=mode, exchange_fuse=fuse)
            l0 = h.launch_count()
            ms = timeit(fn, reps, warm)
            emit(f"{name} [{label}]", ms, byts, launches_per_apply=(h.launch_count() - l0) / (reps + warm))
        h.set_tuning(exchange_mode=0, exchange_fuse=0)

    def timeit(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            allt = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            per_rank[:] = [round(float(v[0]), 4) for v in allt]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        else:
            per_rank[:] = [round(float(t[0]), 4)]
        return float(t[0])

    per_rank = []

    def emit(name, ms, byts=None, flops=None, **kw):
        if rank:
            return
        r = {"config": name, "n_gpus": world, "ms": ms, "ms_per_rank": list(per_rank)}
        if byts is not None:
            r["GBs"] = byts / ms / 1e6
            r["frac_of_aggregate_hbm"] = r["GBs"] / (PEAK * world)
        if flops is not None:
            r["TFLOPs"] = flops / ms / 1e9
        r.update(kw)
        print(json.dumps(r), flush=True)

    def vec(n, seed, off=0, dt=torch.float64):
        t = torch.empty(n, dtype=dt, device=dev)
        if n:
            L.fill_uniform(t, seed, offset=off)
        return t

    # ---- GM: narrow banded x banded materialise n = 2^26, column-partitioned (static halo of A, no communication) ----
    for l in ((1, 8) if want("GM") else ()):
        n = 1 << 26
        q = GbmmColumnShard.plan(n, l, l, l, l, rank, world)
        nb = 2 * l + 1
        A_slab = torch.empty((q.pb - q.pa, nb), dtype=torch.float64, device=dev)
        L.fill_uniform(A_slab, 1, offset=nb * q.pa)
        B_slab = torch.empty((q.c1 - q.c0, nb), dtype=torch.float64, device=dev)
        L.fill_uniform(B_slab, 2, offset=nb * q.c0)
        C_slab = torch.empty((q.c1 - q.c0, 4 * l + 1), dtype=torch.float64, device=dev)
        emit(f"GM gbmm f64 n=2^26 ({l},{l})^2 -> ({2 * l},{2 * l}) register kernel", timeit(lambda: q.local_gbmm(A_slab, B_slab, C_slab), 5, 2),
             8 * n * (2 * nb + 4 * l + 1), 2 * n * nb * nb)
        del A_slab, B_slab, C_slab
        torch.cuda.empty_cache()

    # ---- C2: SymTridiagonal / Bidiagonal mat-vec n = 2^28, row-sharded with ghost exchange ----
    n = 1 << 28
    for nm, has_l, has_u, arrays in ((("SymTridiagonal", True, True, 4), ("Bidiagonal U", False, True, 4)) if want("C2") else ()):
        c0, c1 = ShardedTridiagonal.slices(n, int(has_l), int(has_u), rank, world)
        d = vec(c1 - c0, 1, c0)
        e = vec(c1 - c0 - 1, 2, c0)
        sh = ShardedTridiagonal(n, e if has_l else None, d, e if has_u else None, rank, world)
        p = sh.plan
        xb = p.new_vector(torch.float64, dev)
        p.owned(xb).copy_(vec(p.m_loc, 4, p.r0))
        y = torch.empty(p.m_loc, dtype=torch.float64, device=dev)
        both(f"C2 {nm} mul! f64 n=2^28", lambda: sh.mul_(y, xb), arrays * 8 * n)
        del d, e, sh, xb, y
        torch.cuda.empty_cache()

    # ---- C3: 2-D Laplacian on 8192^2, partitioned by grid columns ----
    if want("C3"):
        g = 8192
        hh = 1.0 / g
        D = L.banded_from_diagonals(g, {0: -2.0 / hh**2, 1: 1.0 / hh**2, -1: 1.0 / hh**2}, device=dev)
        p = RowShard.plan(g, 1, 1, rank, world)
        sh = ShardedKronSum(g, g, 1, 1, D.data[p.c0:p.c1].contiguous(), D, rank, world)
        xb = sh.new_buffer(torch.float64, dev)
        L.fill_uniform(sh.owned(xb).view(-1), 5, offset=p.r0 * g)
        y = torch.empty((p.m_loc, g), dtype=torch.float64, device=dev)
        both("C3 2-D Laplacian kron(D,I)+kron(I,D) 8192x8192", lambda: sh.mul_(y, xb), 16 * g * g, 200, 20)
        del sh, xb, y
        torch.cuda.empty_cache()

    # ---- C4: gbmm l=u=128, n = 2^20, column-partitioned (static halo of A, no communication) ----
    if want("C4"):
        n = 1 << 20
        q = GbmmColumnShard.plan(n, 128, 128, 128, 128, rank, world)
        A_slab = torch.empty((q.pb - q.pa, 257), dtype=torch.float64, device=dev)
        L.fill_uniform(A_slab, 1, offset=257 * q.pa)
        B_slab = torch.empty((q.c1 - q.c0, 257), dtype=torch.float64, device=dev)
        L.fill_uniform(B_slab, 2, offset=257 * q.c0)
        C_slab = torch.empty((q.c1 - q.c0, 513), dtype=torch.float64, device=dev)
        emit("C4 gbmm f64 n=2^20 (128,128)^2 -> (256,256)", timeit(lambda: q.local_gbmm(A_slab, B_slab, C_slab), 3, 1),
             8 * n * (257 + 257 + 513), 2 * n * 257 * 257)
        del A_slab, B_slab, C_slab
        torch.cuda.empty_cache()

    # ---- C5: 4096 fused chains A*B*C, brand(4096,4096,4,4) f32, batch split ----
    if want("C5"):
        batch, nn = 4096, 4096
        b0, b1 = split_batch(batch, world, rank)
        fs = [torch.empty((b1 - b0, nn, 9), dtype=torch.float32, device=dev) for _ in range(3)]
        for i, f in enumerate(fs):
            L.fill_uniform(f, 1 + i, offset=b0 * nn * 9)
        emit("C5 fused chain A*B*C f32 batch=4096 n=4096 (4,4)^3", timeit(lambda: L.chain3_batched(*fs, [4, 4, 4], [4, 4, 4]), 5, 2),
             batch * 4 * nn * (9 + 9 + 9 + 25), batch * 2 * nn * (9 * 9 + 17 * 9))
        del fs
        torch.cuda.empty_cache()

    # ---- M: gbmv n = 2^27, l=u=1 / 8 (the metric's timed region), every transport ----
    for l in ((1, 8) if want("M") else ()):
        n = 1 << 27
        sh = ShardedBanded.brand(n, l, l, rank, world, device=dev, seed=1)
        xb = sh.plan.new_vector(torch.float64, dev)
        L.fill_uniform(xb, 9, offset=sh.plan.c0)
        y = torch.empty(sh.plan.m_loc, dtype=torch.float64, device=dev)
        both(f"M gbmv f64 n=2^27 l=u={l}", lambda: sh.mul_(y, xb), 8 * n * (2 * l + 3), 20, 5)
        # the same slab kernels with NO exchange at all (stale ghosts: timing only) -- what the ranks cost uncoupled
        emit(f"M gbmv f64 n=2^27 l=u={l} [no exchange: timing only]", timeit(lambda: sh.mul_(y, xb, exchange=False), 20, 5), 8 * n * (2 * l + 3))
        del sh, xb, y
        torch.cuda.empty_cache()
    if only is not None and not want("M128"):
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- M: gbmv n = 2^27, l=u=128 (needs >= 2 GPUs) ----
    if world >= 2:
        n = 1 << 27
        sh = ShardedBanded.brand(n, 128, 128, rank, world, device=dev, seed=1)
        xb = sh.plan.new_vector(torch.float64, dev)
        L.fill_uniform(xb, 9, offset=sh.plan.c0)
        y = torch.empty(sh.plan.m_loc, dtype=torch.float64, device=dev)
        both("M gbmv f64 n=2^27 l=u=128", lambda: sh.mul_(y, xb), 8 * n * 259, 5, 2)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
