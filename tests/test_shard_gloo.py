"""N>1 host logic on CPU: world_size-2 (and 3) gloo process groups run the row-sharding plan and
the ghost-cell halo exchange of lbm_b200.shard, with the oracle standing in for the local CUDA
kernel (tests may use the oracle; the product path cannot).  Checks the sharded result equals
the single-process oracle BIT-EXACTLY (same per-row summation order)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_local_gbmv(y, data, m, n, kl, ku, x, alpha, beta):
    from oracle import oracle as O
    A = np.asfortranarray(data.numpy().T)
    out = O.gbmv("N", m, n, kl, ku, alpha, A, x.numpy(), beta, y.numpy(), lda=A.shape[0])
    y.copy_(torch.from_numpy(out))
    return y


def _worker(rank, world, port, n, bands, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "lazybandedmatrices.jl_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from lbm_b200.shard import RowShard, ShardedBanded, sharded_chain_mul_
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        x_full = O.fill_uniform(0x9, n)
        factors = []
        for fi, (l, u) in enumerate(bands):
            lda = l + u + 1
            p = RowShard.plan(n, l, u, rank, world)
            # the rank's column slice of the global brand(n,n,l,u), regenerated from the stream
            slab = O.fill_uniform(0x1 + fi, lda * p.n_loc, offset=lda * p.c0).reshape(p.n_loc, lda)
            factors.append(ShardedBanded(p, torch.from_numpy(slab)))
        # single mul!: y = F_last x
        f = factors[-1]
        xbuf = f.plan.new_vector(torch.float64, "cpu")
        f.plan.owned(xbuf).copy_(torch.from_numpy(x_full[f.plan.r0:f.plan.r1]))
        y = torch.zeros(f.plan.m_loc, dtype=torch.float64)
        f.mul_(y, xbuf, local_gbmv=_oracle_local_gbmv)
        out["single"] = (f.plan.r0, y.numpy().copy())
        y_no = torch.zeros(f.plan.m_loc, dtype=torch.float64)      # non-overlapped path, same answer
        f.mul_(y_no, xbuf, local_gbmv=_oracle_local_gbmv, overlap=False)
        assert torch.equal(y_no, y)
        a, b = f.interior_rows()
        assert 0 <= a <= b <= f.plan.m_loc and (f.plan.left == 0 or a >= min(f.plan.l, f.plan.m_loc))
        assert a % 4 == 0 or a == f.plan.m_loc        # a short last shard has no interior at all
        # ghosts really were filled from the neighbours
        assert np.array_equal(xbuf.numpy(), x_full[f.plan.c0:f.plan.c1])
        # lazy chain: y = F1 (F2 (... x))
        tmp = [factors[len(factors) - 2 - i].plan.new_vector(torch.float64, "cpu") for i in range(len(factors) - 1)]
        y2 = torch.zeros(factors[0].plan.m_loc, dtype=torch.float64)
        sharded_chain_mul_(y2, factors, xbuf, tmp, local_gbmv=_oracle_local_gbmv)
        out["chain"] = (factors[0].plan.r0, y2.numpy().copy())
        q.put((rank, out))
        dist.barrier()
    except BaseException as e:      # report instead of leaving the other ranks (and the test) hanging in a collective
        import traceback
        q.put((rank, {"error": traceback.format_exc() or repr(e)}))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,bands", [
    (2, 1000, [(1, 1), (1, 1)]),
    (2, 4099, [(8, 8), (2, 5), (1, 1)]),
    (3, 2000, [(128, 128), (3, 0)]),
    (2, 64, [(0, 1), (1, 0)]),
    # trailing shards EMPTY or SHORT (n < 4*world / the last shard narrower than the halo): the effective world shrinks,
    # empty ranks have no neighbours, nobody posts an unmatched send (csrc/dist.cu:make_plan, shard.py:effective_world)
    (4, 10, [(1, 1), (2, 1)]),
    (8, 100, [(8, 5), (1, 1)]),
    (3, 5, [(1, 1)]),
])
def test_row_sharded_mul_matches_single_process(world, n, bands, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bands, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    try:
        for _ in range(world):
            rank, out = q.get(timeout=120)
            assert "error" not in out, f"rank {rank}: {out['error']}"
            res[rank] = out
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    # single-process oracle on the full operator
    x = oracle.fill_uniform(0x9, n)
    mats = [np.asfortranarray(oracle.fill_uniform(0x1 + i, (l + u + 1) * n).reshape((l + u + 1, n), order="F"))
            for i, (l, u) in enumerate(bands)]
    l, u = bands[-1]
    want_single = oracle.gbmv("N", n, n, l, u, 1.0, mats[-1], x)
    v = x
    for (l, u), M in zip(reversed(bands), reversed(mats)):
        v = oracle.gbmv("N", n, n, l, u, 1.0, M, v)
    for key, want in (("single", want_single), ("chain", v)):
        got = np.zeros(n)
        covered = np.zeros(n, dtype=bool)
        for r in range(world):
            r0, yy = res[r][key]
            got[r0:r0 + len(yy)] = yy
            covered[r0:r0 + len(yy)] = True
        assert covered.all()
        assert np.array_equal(got, want), key


def test_plan_geometry():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "lazybandedmatrices.jl_b200"))
    from lbm_b200.shard import RowShard, split_rows
    n, l, u, W = 1 << 27, 8, 8, 8
    plans = [RowShard.plan(n, l, u, r, W) for r in range(W)]
    assert plans[0].r0 == 0 and plans[-1].r1 == n
    for a, b in zip(plans[:-1], plans[1:]):
        assert a.r1 == b.r0
    assert plans[0].left == 0 and plans[0].kl_loc == l and plans[0].ku_loc == u
    assert plans[3].left == l and plans[3].kl_loc == 0 and plans[3].ku_loc == l + u
    assert plans[-1].right == 0
    assert all(p.r0 % 4 == 0 for p in plans)
    # uneven split, world that does not divide n
    rr = [split_rows(1003, 3, r) for r in range(3)]
    assert rr[0][0] == 0 and rr[-1][1] == 1003 and all(a[1] == b[0] for a, b in zip(rr[:-1], rr[1:]))
    from lbm_b200.errors import LbmArgumentError
    with pytest.raises(LbmArgumentError):
        RowShard.plan(64, 128, 128, 0, 4)
    # the verdict is a function of (n, l, u, world) only: EVERY rank fails alike, including the empty / short ones
    for r in range(8):
        with pytest.raises(LbmArgumentError):
            RowShard.plan(100, 20, 0, r, 8)
    # n=100, world=8: shards of 16 rows -> 7 ranks own rows (the last one 4 rows), rank 7 owns nothing and has no neighbours
    ps = [RowShard.plan(100, 8, 5, r, 8) for r in range(8)]
    assert [p.m_loc for p in ps] == [16] * 6 + [4, 0] and ps[0].world_eff == 7
    assert [p.has_right for p in ps] == [True] * 6 + [False, False] and [p.has_left for p in ps] == [False] + [True] * 6 + [False]
    assert ps[7].n_loc == 0 and ps[5].right == 4 and ps[6].left == 8 and ps[6].right == 0
    # the C library's plan (used by lbm_?gbmv_sharded) is the same function, bit for bit
    import ctypes as C
    import lbm_b200
    lib = lbm_b200.lib()
    out = (C.c_int64 * 8)()
    for r in range(8):
        assert lib.lbm_shard_plan(100, 20, 0, r, 8, out) == -2
    for (nn, ll, uu, ww) in [(1 << 27, 8, 8, 8), (1003, 1, 1, 3), (4099, 128, 128, 2), (64, 0, 1, 2), (10, 2, 3, 1), (100, 8, 5, 8),
                             (10, 1, 1, 4), (5, 1, 1, 3), (0, 1, 1, 2)]:
        for r in range(ww):
            pl = RowShard.plan(nn, ll, uu, r, ww)
            assert lib.lbm_shard_plan(nn, ll, uu, r, ww, out) == 0
            assert list(out) == [pl.r0, pl.r1, pl.c0, pl.c1, pl.left, pl.right, pl.kl_loc, pl.ku_loc]
    assert lib.lbm_shard_plan(10, 1, 1, 3, 3, out) == -4


def test_gbmm_column_partition_matches_global(oracle):
    """C = A*B column-partitioned over ranks with a static halo of A and
    NO communication -- each rank's sub-product (shifted bandwidths) reproduces its column slice of
    the global product bit-exactly (host logic of lbm_b200.shard.GbmmColumnShard, oracle as kernel)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "lazybandedmatrices.jl_b200"))
    from lbm_b200.shard import GbmmColumnShard, split_batch

    def oracle_gbmm(m, n, k, A, kla, kua, B, klb, kub, Cs, klc, kuc):
        out = oracle.gbmm(m, n, k, np.asfortranarray(A.numpy().T), kla, kua, np.asfortranarray(B.numpy().T), klb, kub,
                          klc=klc, kuc=kuc)
        Cs.copy_(torch.from_numpy(np.ascontiguousarray(out.T)))

    for (n, la, ua, lb, ub, world) in [(200, 1, 1, 1, 1, 2), (333, 8, 3, 2, 5, 4), (1500, 128, 128, 128, 128, 3),
                                       (64, 0, 2, 3, 0, 2)]:
        A = oracle.brand(n, n, la, ua, seed=1)
        B = oracle.brand(n, n, lb, ub, seed=2)
        want = oracle.gbmm(n, n, n, A, la, ua, B, lb, ub)                 # (lc+uc+1) x n
        got = np.zeros_like(want)
        At, Bt = torch.from_numpy(np.ascontiguousarray(A.T)), torch.from_numpy(np.ascontiguousarray(B.T))
        for r in range(world):
            p = GbmmColumnShard.plan(n, la, ua, lb, ub, r, world)
            (kla, kua), (klb, kub), (klc, kuc) = p.sub_bandwidths()
            assert min(kla, kua, klb, kub, klc, kuc) >= 0
            Cs = p.local_gbmm(At[p.pa:p.pb], Bt[p.c0:p.c1], gbmm_fn=oracle_gbmm)
            got[:, p.c0:p.c1] = Cs.numpy().T
        # in-matrix entries identical (same summation order); out-of-matrix slots are zero in both
        assert np.array_equal(oracle.band_to_dense(got, n, la + lb, ua + ub), oracle.band_to_dense(want, n, la + lb, ua + ub))
    assert [split_batch(4096, 8, r) for r in (0, 7)] == [(0, 512), (3584, 4096)]
    assert [split_batch(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]


def _kron_worker(rank, world, port, n1, n2, bands, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "lazybandedmatrices.jl_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from lbm_b200.shard import ShardedKronSum
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        (kla, kua), (klb, kub) = bands
        A = O.brand(n1, n1, kla, kua, seed=3)
        Bn = O.brand(n2, n2, klb, kub, seed=4)
        Bd = O.band_to_dense(Bn, n2, klb, kub)
        x = O.fill_uniform(5, n1 * n2).reshape(n1, n2)          # row c = grid column c

        class _B:                                               # only .l/.u/.data/.lda are consulted on GPU
            pass

        def local_apply(y_own, sh, xbuf, alpha, beta):
            p = sh.plan
            slab = np.asfortranarray(sh.a_slab.numpy().T)
            Aloc = O.band_to_dense(slab, p.m_loc, p.kl_loc, p.ku_loc)      # (r1-r0) x (c1-c0)
            Xext = xbuf.numpy().T                                           # n2 x (c1-c0)
            Y = Xext @ Aloc.T + Bd @ Xext[:, p.left:p.left + p.m_loc]
            y_own.copy_(torch.from_numpy(np.ascontiguousarray((alpha * Y).T)) + beta * y_own)
            return y_own

        from lbm_b200.shard import RowShard
        p = RowShard.plan(n1, kla, kua, rank, world)
        sh = ShardedKronSum(n1, n2, kla, kua, torch.from_numpy(np.ascontiguousarray(A.T[p.c0:p.c1])), _B(), rank, world)
        xbuf = sh.new_buffer(torch.float64, "cpu")
        sh.owned(xbuf).copy_(torch.from_numpy(x[p.r0:p.r1]))
        y = torch.zeros((p.m_loc, n2), dtype=torch.float64)
        sh.mul_(y, xbuf, local_apply=local_apply)
        assert np.array_equal(xbuf.numpy(), x[p.c0:p.c1])       # ghost columns filled from the neighbours
        q.put((rank, p.r0, y.numpy().copy()))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n1,n2,bands", [(2, 40, 16, ((1, 1), (1, 1))), (3, 50, 12, ((2, 1), (0, 2)))])
def test_kronsum_column_partition_gloo(world, n1, n2, bands, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_kron_worker, args=(r, world, port, n1, n2, bands, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = np.zeros((n1, n2))
    for _ in range(world):
        rank, r0, yy = q.get(timeout=120)
        got[r0:r0 + yy.shape[0]] = yy
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (kla, kua), (klb, kub) = bands
    A = oracle.brand(n1, n1, kla, kua, seed=3)
    Bn = oracle.brand(n2, n2, klb, kub, seed=4)
    x = oracle.fill_uniform(5, n1 * n2)
    want = oracle.kronsum2d_mv(n1, n2, A, kla, kua, Bn, klb, kub, x).reshape(n1, n2)
    assert np.max(np.abs(got - want)) <= 4 * np.finfo(np.float64).eps * (kla + kua + klb + kub + 2)


def _tri_worker(rank, world, port, n, kind, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "lazybandedmatrices.jl_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from lbm_b200.shard import RowShard, ShardedTridiagonal
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d, dl, du, x = O.fill_uniform(1, n), O.fill_uniform(2, n - 1), O.fill_uniform(3, n - 1), O.fill_uniform(4, n)
        if kind == "sym":
            dl = du
        has_l, has_u = kind != "biU", kind != "biL"
        c0, c1 = ShardedTridiagonal.slices(n, int(has_l), int(has_u), rank, world)
        t = lambda a: torch.from_numpy(a.copy())
        sh = ShardedTridiagonal(n, t(dl[c0:c1 - 1]) if has_l else None, t(d[c0:c1]), t(du[c0:c1 - 1]) if has_u else None,
                                rank, world)
        p = sh.plan
        xbuf = p.new_vector(torch.float64, "cpu")
        p.owned(xbuf).copy_(t(x[p.r0:p.r1]))

        def local_apply(yb, nl, a, dd, b, xb, alpha, beta):
            out = O.tridiag_mv(nl, None if a is None else a.numpy(), dd.numpy(), None if b is None else b.numpy(),
                               xb.numpy(), alpha, beta, yb.numpy())
            yb.copy_(torch.from_numpy(out))

        y = torch.zeros(p.m_loc, dtype=torch.float64)
        sh.mul_(y, xbuf, local_apply=local_apply)
        q.put((rank, p.r0, y.numpy().copy()))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["tri", "sym", "biU", "biL"])
def test_sharded_tridiagonal_gloo(kind, oracle):
    world, n = 3, 1003
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tri_worker, args=(r, world, port, n, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = np.zeros(n)
    for _ in range(world):
        rank, r0, yy = q.get(timeout=120)
        got[r0:r0 + len(yy)] = yy
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d, dl, du, x = oracle.fill_uniform(1, n), oracle.fill_uniform(2, n - 1), oracle.fill_uniform(3, n - 1), oracle.fill_uniform(4, n)
    if kind == "sym":
        dl = du
    want = oracle.tridiag_mv(n, dl if kind != "biU" else None, d, du if kind != "biL" else None, x)
    assert np.array_equal(got, want)
