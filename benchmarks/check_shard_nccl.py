#!/usr/bin/env python
"""N>1 parity on real GPUs (run under torchrun): row-sharded mul! and a lazy chain through the
CUDA kernels + NCCL ghost exchange vs the single-process oracle.  Prints OK / raises."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import lbm_b200 as L  # noqa: E402
from lbm_b200.shard import (RowShard, ShardedBanded, ShardedKronSum, ShardedTridiagonal, GbmmColumnShard, init_comm,  # noqa: E402
                            sharded_chain_mul_)
from oracle import oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eps = np.finfo(np.float64).eps
    h = L.handle_for(local)
    _check(rank, world, dev, eps)                       # torch.distributed P2P path (no library communicator yet)
    init_comm(local)                                    # library communicator + peer-memory mailboxes
    assert h.comm_world == world
    peer = bool(L._lib.lib().lbm_peer_active(h.ptr))
    # peer mailboxes with the exchange FUSED into the compute launch (default), the same mailboxes through the two-kernel
    # exchange (exchange_fuse = 2), then in-library ncclSend/ncclRecv
    for mode, fuse in (((0, 0), (0, 2), (1, 0)) if peer else ((1, 0),)):
        h.set_tuning(exchange_mode=mode, exchange_fuse=fuse)
        _check(rank, world, dev, eps)
        _check_kron_gbmm(rank, world, dev, eps)
        _check_epochs(rank, world, dev, eps)
        _check_short_shards(rank, world, dev, eps)
        _check_tridiag(rank, world, dev, eps)
    h.set_tuning(exchange_mode=0, exchange_fuse=0)
    if peer:
        _check_launch_counts(rank, world, dev, h)
    dist.barrier()
    if rank == 0:
        print(f"OK sharded mul!/chain/kron/tridiag on {world} GPUs match the oracle (torch P2P, "
              f"{'peer-memory mailboxes, ' if peer else 'NO peer access: '}in-library NCCL)")
    dist.destroy_process_group()


def _check_launch_counts(rank, world, dev, h):
    """with the fused exchange a sharded apply is ONE launch per rank (was 4.75)."""
    for (n, l, u) in ((1 << 22, 1, 1), (1 << 22, 8, 8)):
        f = ShardedBanded.brand(n, l, u, rank, world, device=dev, seed=3)
        xb = f.plan.new_vector(torch.float64, dev)
        y = torch.empty(f.plan.m_loc, dtype=torch.float64, device=dev)
        f.mul_(y, xb)
        l0 = h.launch_count()
        for _ in range(10):
            f.mul_(y, xb)
        assert h.launch_count() - l0 == 10, ("launches per apply", (h.launch_count() - l0) / 10)
    g = 2048
    D = L.banded_from_diagonals(g, {0: -2.0, 1: 1.0, -1: 1.0}, device=dev)
    p = RowShard.plan(g, 1, 1, rank, world)
    sh = ShardedKronSum(g, g, 1, 1, D.data[p.c0:p.c1].contiguous(), D, rank, world)
    xb = sh.new_buffer(torch.float64, dev)
    y = torch.empty((p.m_loc, g), dtype=torch.float64, device=dev)
    sh.mul_(y, xb)
    l0 = h.launch_count()
    for _ in range(10):
        sh.mul_(y, xb)
    assert h.launch_count() - l0 == 10, ("kronsum launches per apply", (h.launch_count() - l0) / 10)
    torch.cuda.synchronize()


def _check_short_shards(rank, world, dev, eps):
    """trailing shards empty or shorter than the halo (n < 4*world, ...) must neither hang nor fault: the
    effective world shrinks, empty ranks return at once, everyone agrees."""
    for (n, l, u) in ((100, 8, 5), (4 * world - 2, 1, 1), (16 * world - 13, 2, 3)):
        try:
            f = ShardedBanded.brand(n, l, u, rank, world, device=dev, seed=0x33)
        except L.LbmArgumentError:
            continue        # shards narrower than the halo: every rank raises alike (pure function of n, l, u, world)
        M = np.asfortranarray(O.fill_uniform(0x33, (l + u + 1) * n).reshape((l + u + 1, n), order="F"))
        x_full = O.fill_uniform(0x9, n)
        want = O.gbmv("N", n, n, l, u, 1.0, M, x_full)[f.plan.r0:f.plan.r1]
        xbuf = f.plan.new_vector(torch.float64, dev)
        f.plan.owned(xbuf).copy_(torch.from_numpy(x_full[f.plan.r0:f.plan.r1]).to(dev))
        for k in range(6):
            y = torch.full((f.plan.m_loc,), float("nan"), dtype=torch.float64, device=dev)
            f.mul_(y, xbuf, overlap=bool(k % 2 == 0))
            if f.plan.m_loc:
                err = np.max(np.abs(y.cpu().numpy() - want))
                assert err <= 4 * eps * (l + u + 1), ("short shards", n, l, u, k, err)
    torch.cuda.synchronize()


def _check_tridiag(rank, world, dev, eps):
    """C2 path: one C call per apply (lbm_?tridiag_mv_sharded), all three kinds, x changing between applies."""
    n = 1 << 20
    d_full, e_full, x_full = O.fill_uniform(1, n), O.fill_uniform(2, n - 1), O.fill_uniform(4, n)
    for has_l, has_u in ((1, 1), (0, 1), (1, 0)):
        c0, c1 = ShardedTridiagonal.slices(n, has_l, has_u, rank, world)
        d = torch.from_numpy(d_full[c0:c1]).to(dev)
        e = torch.from_numpy(e_full[c0:c1 - 1]).to(dev)
        sh = ShardedTridiagonal(n, e if has_l else None, d, e if has_u else None, rank, world)
        p = sh.plan
        want = O.tridiag_mv(n, e_full if has_l else None, d_full, e_full if has_u else None, x_full)[p.r0:p.r1]
        xb = p.new_vector(torch.float64, dev)
        for k in range(8):
            p.owned(xb).copy_(torch.from_numpy(x_full[p.r0:p.r1] * float(2 ** (k % 3))).to(dev))
            y = torch.full((p.m_loc,), float("nan"), dtype=torch.float64, device=dev)
            sh.mul_(y, xb, overlap=bool(k % 2 == 0))
            err = np.max(np.abs(y.cpu().numpy() / float(2 ** (k % 3)) - want))
            assert err <= 4 * eps * 3, ("tridiag", has_l, has_u, k, err)


def _check_epochs(rank, world, dev, eps):
    """Back-to-back applies with x changing every time (scaled by exact powers of two): every epoch must see the
    neighbours' CURRENT boundary values -- exercises the mailbox double-buffering and the flag handshake."""
    n, l, u = 300007, 8, 5
    f = ShardedBanded.brand(n, l, u, rank, world, device=dev, seed=0x21)
    M = np.asfortranarray(O.fill_uniform(0x21, (l + u + 1) * n).reshape((l + u + 1, n), order="F"))
    x_full = O.fill_uniform(0x9, n)
    want = O.gbmv("N", n, n, l, u, 1.0, M, x_full)[f.plan.r0:f.plan.r1]
    x_own = torch.from_numpy(x_full[f.plan.r0:f.plan.r1]).to(dev)
    xbuf = f.plan.new_vector(torch.float64, dev)
    ys = []
    for k in range(40):
        f.plan.owned(xbuf).copy_(x_own * float(2 ** (k % 7)))
        y = torch.empty(f.plan.m_loc, dtype=torch.float64, device=dev)
        f.mul_(y, xbuf, overlap=bool(k % 2))
        ys.append(y)
    for k, y in enumerate(ys):
        err = np.max(np.abs(y.cpu().numpy() / float(2 ** (k % 7)) - want))
        assert err <= 4 * eps * (l + u + 1), ("epoch", k, err)


def _check_kron_gbmm(rank, world, dev, eps):
    """C3 (grid-column partition, in-library ghost-column exchange) and C4 (column-partitioned gbmm)."""
    for (n1, n2, (kla, kua), (klb, kub)) in ((4096, 2048, (1, 1), (1, 1)), (1000, 516, (2, 2), (1, 1)), (300, 70, (1, 2), (3, 0))):
        A = O.brand(n1, n1, kla, kua, seed=3)
        Bn = O.brand(n2, n2, klb, kub, seed=4)
        x = O.fill_uniform(5, n1 * n2)
        want = O.kronsum2d_mv(n1, n2, A, kla, kua, Bn, klb, kub, x, 1.5, 0.0).reshape(n1, n2)
        p = RowShard.plan(n1, kla, kua, rank, world)
        Bg = L.BandedMatrix(torch.from_numpy(np.ascontiguousarray(Bn.T)).to(dev), n2, klb, kub)
        sh = ShardedKronSum(n1, n2, kla, kua, torch.from_numpy(np.ascontiguousarray(A.T[p.c0:p.c1])).to(dev), Bg, rank, world)
        xbuf = sh.new_buffer(torch.float64, dev)
        sh.owned(xbuf).copy_(torch.from_numpy(x.reshape(n1, n2)[p.r0:p.r1]))
        for overlap in (True, False):
            y = torch.full((p.m_loc, n2), float("nan"), dtype=torch.float64, device=dev)
            sh.mul_(y, xbuf, alpha=1.5, overlap=overlap)
            err = np.max(np.abs(y.cpu().numpy() - want[p.r0:p.r1]))
            assert err <= 4 * eps * 1.5 * (kla + kua + klb + kub + 2), ("kronsum", n1, n2, overlap, err)
    for kind in ("tri", "sym", "biU", "biL"):                       # C2: row-sharded structured mat-vec
        n = 400003
        d, dl, du, x = O.fill_uniform(1, n), O.fill_uniform(2, n - 1), O.fill_uniform(3, n - 1), O.fill_uniform(4, n)
        if kind == "sym":
            dl = du
        has_l, has_u = kind != "biU", kind != "biL"
        c0, c1 = ShardedTridiagonal.slices(n, int(has_l), int(has_u), rank, world)
        t = lambda a: torch.from_numpy(a.copy()).to(dev)  # noqa: E731
        sh = ShardedTridiagonal(n, t(dl[c0:c1 - 1]) if has_l else None, t(d[c0:c1]), t(du[c0:c1 - 1]) if has_u else None, rank, world)
        xbuf = sh.plan.new_vector(torch.float64, dev)
        sh.plan.owned(xbuf).copy_(t(x[sh.plan.r0:sh.plan.r1]))
        y = torch.full((sh.plan.m_loc,), float("nan"), dtype=torch.float64, device=dev)
        sh.mul_(y, xbuf)
        want = O.tridiag_mv(n, dl if has_l else None, d, du if has_u else None, x)
        err = np.max(np.abs(y.cpu().numpy() - want[sh.plan.r0:sh.plan.r1]))
        assert err <= 4 * eps * 3, ("sharded " + kind, err)
    for (n, la, ua, lb, ub) in ((5000, 1, 1, 1, 1), (3000, 128, 128, 128, 128), (2000, 8, 3, 2, 5)):
        A, B = O.brand(n, n, la, ua, seed=1), O.brand(n, n, lb, ub, seed=2)
        want = O.gbmm(n, n, n, A, la, ua, B, lb, ub)
        p = GbmmColumnShard.plan(n, la, ua, lb, ub, rank, world)
        Cs = p.local_gbmm(torch.from_numpy(np.ascontiguousarray(A.T[p.pa:p.pb])).to(dev),
                          torch.from_numpy(np.ascontiguousarray(B.T[p.c0:p.c1])).to(dev))
        got = np.zeros_like(want)
        got[:, p.c0:p.c1] = Cs.cpu().numpy().T
        ref = np.zeros_like(want)
        ref[:, p.c0:p.c1] = want[:, p.c0:p.c1]
        dg, dw = O.band_to_dense(got, n, la + lb, ua + ub), O.band_to_dense(ref, n, la + lb, ua + ub)
        assert np.max(np.abs(dg - dw)) <= 4 * eps * min(la + ua + 1, lb + ub + 1), ("gbmm shard", n, la)


def _check(rank, world, dev, eps):
    for n, bands in ((200003, [(1, 1), (1, 1)]), (100000, [(8, 8), (2, 5), (1, 1)]), (50000, [(128, 128), (3, 0)])):
        x_full = O.fill_uniform(0x9, n)
        factors = [ShardedBanded.brand(n, l, u, rank, world, device=dev, seed=0x1 + i) for i, (l, u) in enumerate(bands)]
        f = factors[-1]
        xbuf = f.plan.new_vector(torch.float64, dev)
        f.plan.owned(xbuf).copy_(torch.from_numpy(x_full[f.plan.r0:f.plan.r1]))
        mats = [np.asfortranarray(O.fill_uniform(0x1 + i, (l + u + 1) * n).reshape((l + u + 1, n), order="F"))
                for i, (l, u) in enumerate(bands)]
        l, u = bands[-1]
        want = O.gbmv("N", n, n, l, u, 1.0, mats[-1], x_full)
        for overlap in (True, False):
            y = torch.full((f.plan.m_loc,), float("nan"), dtype=torch.float64, device=dev)
            f.mul_(y, xbuf, overlap=overlap)
            err = np.max(np.abs(y.cpu().numpy() - want[f.plan.r0:f.plan.r1]))
            assert err <= 4 * eps * (l + u + 1), (n, bands, overlap, err)
        tmp = [factors[len(factors) - 2 - i].plan.new_vector(torch.float64, dev) for i in range(len(factors) - 1)]
        y2 = torch.empty(factors[0].plan.m_loc, dtype=torch.float64, device=dev)
        sharded_chain_mul_(y2, factors, xbuf, tmp)
        v = x_full
        scale = 1.0
        for (l, u), M in zip(reversed(bands), reversed(mats)):
            v = O.gbmv("N", n, n, l, u, 1.0, M, v)
            scale *= (l + u + 1)
        err = np.max(np.abs(y2.cpu().numpy() - v[factors[0].plan.r0:factors[0].plan.r1]))
        assert err <= 4 * eps * scale * len(bands), (n, bands, "chain", err)


if __name__ == "__main__":
    main()
