"""Row-sharding of banded `mul!` across the GPUs of one box: one process per GPU
(`torch.distributed`, NCCL over NVLink 5 / NVSwitch; gloo on CPU for the host-logic tests).

Output rows are independent given an x window: rank r owns rows [r0, r1) and needs
x[r0-l, r1+u).  In diagonal-major storage the rows [r0,r1) of A live in the COLUMN slice
[c0,c1) = [max(0,r0-l), min(n,r1+u)) of the global `data` array, and that slice, taken as is
(same lda, same band-row index u+i-j), is a (r1-r0) x (c1-c0) banded matrix with bandwidths
(l - left, u + left), left = r0 - c0.  So a shard is an ordinary rectangular lbm_?gbmv on a
ghost-cell vector [left ghosts | owned | right ghosts]; the only communication is the
nearest-neighbour exchange that fills the max(l,u)-wide ghosts.  There is no
all-reduce / all-gather on the path, and no replicated compute.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.distributed as dist

from .errors import LbmArgumentError


def init_comm(device: int, group=None, peer: bool = True, slot_bytes: int = 1 << 20):
    """Collective: create the library's own NCCL communicator for this process group, so that a
    sharded mul! (ghost exchange + interior + boundary kernels) is ONE C call
    (lbm_?gbmv_sharded).  Rank 0 draws the NCCL unique id, torch.distributed broadcasts its
    128 bytes, every rank joins.  With `peer=True` (default) the ghost transport is then switched to peer-memory
    mailboxes (connect_peers); `lbm_set_tuning("exchange_mode", 1)` goes back to NCCL send/recv at run time."""
    import ctypes as C

    from . import _lib
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    h = _lib.handle_for(device)
    if world == 1:
        return h
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        rc = _lib.lib().lbm_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc:
            raise RuntimeError(f"lbm_comm_unique_id failed ({rc}): is libnccl.so.2 loadable?")
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        t = t.cuda(device)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    raw = bytes(t.cpu().tolist())
    idb = (C.c_ubyte * 128).from_buffer_copy(raw)
    h.check(_lib.lib().lbm_comm_init(h.ptr, C.cast(idb, C.c_void_p), rank, world), "lbm_comm_init")
    h.comm_world = world
    if peer and dist.get_backend(group) == "nccl":
        connect_peers(h, device, group, slot_bytes)
    return h


def connect_peers(h, device: int, group=None, slot_bytes: int = 1 << 20) -> bool:
    """Collective: switch the library's ghost exchange to peer-memory mailboxes (NVLink peer stores + flag
    handshakes, lbm_peer_alloc / lbm_peer_connect).  Each rank exports the CUDA IPC handle of its mailbox block,
    the 64-byte handles are all-gathered over the process group, and every rank maps its two neighbours' blocks.
    If ANY rank cannot set it up (no peer access between the two GPUs), every rank stays on NCCL send/recv.
    `slot_bytes`: largest halo (per side) the mailboxes carry -- 1 MiB covers 131072 Float64 ghosts, i.e. sixteen
    8192-point grid columns of the Kronecker-sum apply."""
    import ctypes as C

    from . import _lib
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lib = _lib.lib()
    buf = (C.c_ubyte * 64)()
    ok = lib.lbm_peer_alloc(h.ptr, slot_bytes, C.cast(buf, C.c_void_p)) == 0
    mine = torch.tensor(list(buf) + [1 if ok else 0], dtype=torch.uint8).cuda(device)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    allh = [bytes(t.cpu().tolist()) for t in allh]
    if not all(b[64] == 1 for b in allh):
        lib.lbm_peer_disconnect(h.ptr)
        return False

    def hbuf(r):
        return C.cast((C.c_ubyte * 64).from_buffer_copy(allh[r][:64]), C.c_void_p) if 0 <= r < world else None

    left, right = hbuf(rank - 1), hbuf(rank + 1)
    rc = lib.lbm_peer_connect(h.ptr, left, right)
    flag = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32).cuda(device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) != 1:
        lib.lbm_peer_disconnect(h.ptr)
        return False
    return True


def split_rows(n: int, world: int, rank: int, align: int = 4):
    """Contiguous near-equal row blocks whose starts are multiples of `align` (keeps every
    shard's slab and vectors 16-byte aligned for the bulk-TMA path)."""
    per = rows_per_shard(n, world, align)
    r0 = min(n, rank * per)
    r1 = min(n, r0 + per)
    return r0, r1


def rows_per_shard(n: int, world: int, align: int = 4) -> int:
    per = -(-n // world)
    return max(align, -(-per // align) * align)


def effective_world(n: int, world: int, align: int = 4) -> int:
    """Ranks that own rows.  Rounding the shard size up can leave trailing ranks EMPTY (n=100, world=8: rank 6 has four
    rows, rank 7 none); those ranks have no neighbours and take no part in the exchange (csrc/dist.cu:make_plan)."""
    return max(1, -(-n // rows_per_shard(n, world, align)))


@dataclass
class RowShard:
    n: int          # global size (square n x n operator)
    l: int
    u: int
    rank: int
    world: int
    r0: int
    r1: int
    c0: int
    c1: int

    @classmethod
    def plan(cls, n: int, l: int, u: int, rank: int, world: int) -> "RowShard":
        r0, r1 = split_rows(n, world, rank)
        # validated against EVERY rank's shard (a pure function of n, l, u, world), so all ranks fail alike and none
        # returns while a neighbour still waits for it; a short LAST shard is fine (the ghosts it serves are clipped at n)
        if effective_world(n, world) > 1 and rows_per_shard(n, world) < max(l, u):
            raise LbmArgumentError(
                f"shards of {rows_per_shard(n, world)} rows are narrower than the halo max(l,u)={max(l, u)}: use fewer ranks")
        if r1 == r0:      # beyond the effective world: owns nothing, stores nothing, has no neighbours
            return cls(n, l, u, rank, world, r0, r1, r0, r0)
        return cls(n, l, u, rank, world, r0, r1, max(0, r0 - l), min(n, r1 + u))

    # geometry ------------------------------------------------------------------------
    @property
    def m_loc(self):
        return self.r1 - self.r0

    @property
    def n_loc(self):
        return self.c1 - self.c0

    @property
    def left(self):     # ghost entries before the owned block
        return self.r0 - self.c0

    @property
    def right(self):    # ghost entries after it
        return self.c1 - self.r1

    @property
    def world_eff(self):
        return effective_world(self.n, self.world)

    @property
    def has_left(self):
        return 0 < self.rank < self.world_eff

    @property
    def has_right(self):
        return self.rank < self.world_eff - 1

    @property
    def kl_loc(self):
        return self.l - self.left

    @property
    def ku_loc(self):
        return self.u + self.left

    def owned(self, xbuf: torch.Tensor) -> torch.Tensor:
        """View of the owned block inside a ghost-cell vector."""
        return xbuf[self.left: self.left + self.m_loc]

    def new_vector(self, dtype, device) -> torch.Tensor:
        return torch.zeros(self.n_loc, dtype=dtype, device=device)

    # communication ------------------------------------------------------------------
    def exchange_ops(self, xbuf: torch.Tensor, group=None):
        """P2P ops that fill this rank's ghosts and serve the neighbours' ghosts."""
        ops = []
        own = self.owned(xbuf)
        if self.has_left:
            # previous rank's right ghost = my first u owned entries (clipped at n by construction)
            cnt = min(self.u, self.m_loc)
            if cnt:
                ops.append(dist.P2POp(dist.isend, own[:cnt], self.rank - 1, group))
            if self.left:
                ops.append(dist.P2POp(dist.irecv, xbuf[: self.left], self.rank - 1, group))
        if self.has_right:
            cnt = min(self.l, self.m_loc)
            if cnt:
                ops.append(dist.P2POp(dist.isend, own[self.m_loc - cnt:], self.rank + 1, group))
            if self.right:
                ops.append(dist.P2POp(dist.irecv, xbuf[self.left + self.m_loc:], self.rank + 1, group))
        return ops

    def exchange(self, xbuf: torch.Tensor, group=None):
        ops = self.exchange_ops(xbuf, group)
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


class ShardedBanded:
    """This rank's slab of a global n x n BandedMatrix (bandwidths (l,u)), plus the plan."""

    def __init__(self, plan: RowShard, data_loc: torch.Tensor):
        lda = plan.l + plan.u + 1
        if tuple(data_loc.shape) != (plan.n_loc, lda):
            raise LbmArgumentError(f"local slab must be ({plan.n_loc}, {lda}), got {tuple(data_loc.shape)}")
        self.plan = plan
        self.data = data_loc.contiguous()

    @classmethod
    def brand(cls, n, l, u, rank, world, dtype=torch.float64, device="cuda", seed=1):
        """The rank's slice of `brand(n,n,l,u)` generated in place from the counter-based
        stream (bit-identical to slicing the global matrix; no transfer)."""
        from .banded import fill_uniform
        p = RowShard.plan(n, l, u, rank, world)
        lda = l + u + 1
        data = torch.empty((p.n_loc, lda), dtype=dtype, device=device)
        if data.numel():
            fill_uniform(data, seed, offset=lda * p.c0)
        return cls(p, data)

    def local_matrix(self):
        from .banded import BandedMatrix
        p = self.plan
        return BandedMatrix(self.data, p.m_loc, p.kl_loc, p.ku_loc)

    # ---- row sub-ranges of the slab are again rectangular banded matrices -----------------
    def _sub(self, a: int, b: int):
        """Rows [a,b) of the local matrix as (data_view, m, n, kl, ku, col0): the column slice
        [col0, col1) of the slab with bandwidths shifted by the rows skipped."""
        p = self.plan
        c0 = max(0, a - p.kl_loc)
        c1 = min(p.n_loc, b + p.ku_loc)
        shift = a - c0
        return self.data[c0:c1], b - a, c1 - c0, p.kl_loc - shift, p.ku_loc + shift, c0

    def _apply_rows(self, y_own, xbuf, a, b, alpha, beta, local_gbmv):
        if b <= a:
            return
        data, m, n, kl, ku, c0 = self._sub(a, b)
        if local_gbmv is None:
            from .banded import BandedMatrix, gbmv_
            gbmv_(y_own[a:b], BandedMatrix(data, m, kl, ku), xbuf[c0:c0 + n], alpha, beta, "N")
        else:
            local_gbmv(y_own[a:b], data, m, n, kl, ku, xbuf[c0:c0 + n], alpha, beta)

    def interior_rows(self, align: int = 4):
        """Local rows whose x window lies entirely in the owned block (no ghost needed); the
        start is rounded up to `align` so the interior sub-slab keeps the 16-byte alignment the
        bulk-TMA kernel wants."""
        p = self.plan
        a = 0 if p.left == 0 else -(-p.l // align) * align
        b = p.m_loc if p.right == 0 else p.m_loc - p.u
        a = min(a, p.m_loc)
        b = max(a, b)
        return a, b

    def mul_(self, y_own: torch.Tensor, xbuf: torch.Tensor, alpha=1.0, beta=0.0, group=None,
             local_gbmv: Optional[Callable] = None, exchange: bool = True, overlap: bool = True):
        """y_own <- alpha * A[r0:r1, :] x + beta * y_own; `xbuf` is the ghost-cell vector whose
        owned block holds x[r0:r1].

        With `overlap` the ghost exchange (NCCL send/recv on the communicator's stream) runs
        while the interior rows -- all but the first ~l and last u -- are computed; the few
        boundary rows follow once the ghosts have landed.  `local_gbmv` exists so this host logic
        can be exercised on CPU (gloo) with the oracle in tests; the product path is the CUDA
        kernel."""
        p = self.plan
        if not (exchange and p.world > 1):
            self._apply_rows(y_own, xbuf, 0, p.m_loc, alpha, beta, local_gbmv)
            return y_own
        if local_gbmv is None and group is None and xbuf.is_cuda:
            # product fast path: exchange + kernels inside the library (needs init_comm())
            from . import _lib
            from .banded import _ft, _pfx, _ptr
            # the prepared call (bound function + converted arguments) is cached per operand set: at 8 GPUs an apply is
            # ~0.1 ms and the host side of the call must stay in the noise
            key = (y_own.data_ptr(), xbuf.data_ptr(), float(alpha), float(beta), bool(overlap))
            cached = getattr(self, "_prepared", None)
            if cached is not None and cached[0] == key:
                h, fn, args = cached[1]
                h.use_torch_stream()
                rc = fn(*args)
                if rc:
                    h.check(rc, "lbm_gbmv_sharded")
                return y_own
            h = _lib.handle_of(self.data)
            if h.comm_world == p.world:
                fn = getattr(_lib.lib(), f"lbm_{_pfx(self.data.dtype)}gbmv_sharded")
                ft = _ft(self.data.dtype)
                assert xbuf.is_contiguous() and y_own.is_contiguous() and xbuf.shape[0] == p.n_loc
                args = (h.ptr, p.n, p.l, p.u, ft(alpha), _ptr(self.data), p.l + p.u + 1, _ptr(xbuf), ft(beta), _ptr(y_own),
                        1 if overlap else 0)
                self._prepared = (key, (h, fn, args))
                h.check(fn(*args), "lbm_gbmv_sharded")
                return y_own
        if not overlap:
            p.exchange(xbuf, group)
            self._apply_rows(y_own, xbuf, 0, p.m_loc, alpha, beta, local_gbmv)
            return y_own
        ops = p.exchange_ops(xbuf, group)
        works = dist.batch_isend_irecv(ops) if ops else []
        a, b = self.interior_rows()
        self._apply_rows(y_own, xbuf, a, b, alpha, beta, local_gbmv)       # overlaps the exchange
        for w in works:
            w.wait()
        self._apply_rows(y_own, xbuf, 0, a, alpha, beta, local_gbmv)       # needs the left ghosts
        self._apply_rows(y_own, xbuf, b, p.m_loc, alpha, beta, local_gbmv)  # needs the right ghosts
        return y_own


def sharded_chain_mul_(y_own, factors, xbuf, tmp_bufs, group=None, local_gbmv=None):
    """`mul!(y, ApplyMatrix(*, F1, ..., Fk), x)` row-sharded: right-to-left; stage output i is
    written into the owned block of a ghost-cell vector laid out for the factor that consumes
    it, whose `mul_` then exchanges that intermediate's halo (width max(l,u) of the consumer)."""
    if local_gbmv is None and group is None and xbuf.is_cuda and len(factors) >= 1:
        # product path: ONE C call (lbm_?gbmv_chain_sharded) -- stages, intermediate halos and temporaries inside the library
        import ctypes as C

        from . import _lib
        from .banded import _ft, _pfx, _ptr
        h = _lib.handle_of(xbuf)
        p0 = factors[0].plan
        if p0.world == 1 or h.comm_world == p0.world:
            nf = len(factors)
            I64, VP = C.c_int64 * nf, C.c_void_p * nf
            fn = getattr(_lib.lib(), f"lbm_{_pfx(xbuf.dtype)}gbmv_chain_sharded")
            ft = _ft(xbuf.dtype)
            h.check(fn(h.ptr, nf, p0.n, I64(*[f.plan.l for f in factors]), I64(*[f.plan.u for f in factors]),
                       VP(*[f.data.data_ptr() for f in factors]), I64(*[f.plan.l + f.plan.u + 1 for f in factors]),
                       ft(1.0), _ptr(xbuf), ft(0.0), _ptr(y_own)), "lbm_gbmv_chain_sharded")
            return y_own
    order = list(reversed(factors))          # F_k ... F_1
    v = xbuf                                 # ghost layout of F_k
    for idx, f in enumerate(order[:-1]):
        consumer = order[idx + 1]
        t = tmp_bufs[idx]                    # consumer.plan.new_vector(...)
        f.mul_(consumer.plan.owned(t), v, group=group, local_gbmv=local_gbmv)
        v = t
    return order[-1].mul_(y_own, v, group=group, local_gbmv=local_gbmv)


# ------------------------------------------------------------------------------------------------
# banded x banded materialise, column-partitioned: no communication
# ------------------------------------------------------------------------------------------------
@dataclass
class GbmmColumnShard:
    """Rank r owns the COLUMNS [c0,c1) of C = A*B (n x n, bands (la,ua) x (lb,ub)).  It needs
    B[:, c0:c1] and the columns [pa,pb) = [c0-ub, c1+lb) of A (a one-time, static halo).  Taken as
    they lie in the diagonal-major arrays (same ld, same band-row index), those column slices ARE
    banded matrices of a smaller product with shifted bandwidths:
        rows   [ra, rb) = [c0-uc, c1+lc) ∩ [0,n)        (uc = ua+ub, lc = la+lb)
        A_sub  (rb-ra) x (pb-pa),  bandwidths (la - (ra-pa), ua + (ra-pa))
        B_sub  (pb-pa) x (c1-c0),  bandwidths (lb - (pa-c0), ub + (pa-c0))
        C_sub  (rb-ra) x (c1-c0),  bandwidths (lc - (ra-c0), uc + (ra-c0))
    so the local work is one ordinary lbm_?gbmm call and C's column slice is written in place."""
    n: int
    la: int
    ua: int
    lb: int
    ub: int
    rank: int
    world: int
    c0: int
    c1: int

    @classmethod
    def plan(cls, n, la, ua, lb, ub, rank, world):
        c0, c1 = split_rows(n, world, rank)
        return cls(n, la, ua, lb, ub, rank, world, c0, c1)

    @property
    def lc(self):
        return self.la + self.lb

    @property
    def uc(self):
        return self.ua + self.ub

    @property
    def pa(self):
        return max(0, self.c0 - self.ub)

    @property
    def pb(self):
        return min(self.n, self.c1 + self.lb)

    @property
    def ra(self):
        return max(0, self.c0 - self.uc)

    @property
    def rb(self):
        return min(self.n, self.c1 + self.lc)

    def sub_bandwidths(self):
        """((klA,kuA), (klB,kuB), (klC,kuC)) of the local sub-product."""
        sa, sb, sc = self.ra - self.pa, self.pa - self.c0, self.ra - self.c0
        return ((self.la - sa, self.ua + sa), (self.lb - sb, self.ub + sb), (self.lc - sc, self.uc + sc))

    def local_gbmm(self, A_slab: torch.Tensor, B_slab: torch.Tensor, C_slab: torch.Tensor = None, gbmm_fn=None):
        """A_slab = data_A[pa:pb] ((pb-pa) x (la+ua+1)), B_slab = data_B[c0:c1]; returns / fills
        C_slab = data_C[c0:c1] ((c1-c0) x (lc+uc+1))."""
        (kla, kua), (klb, kub), (klc, kuc) = self.sub_bandwidths()
        m_loc, k_loc, n_loc = self.rb - self.ra, self.pb - self.pa, self.c1 - self.c0
        if C_slab is None:
            C_slab = torch.empty((n_loc, self.lc + self.uc + 1), dtype=A_slab.dtype, device=A_slab.device)
        if gbmm_fn is None:
            from .banded import BandedMatrix, gbmm
            gbmm(BandedMatrix(A_slab, m_loc, kla, kua), BandedMatrix(B_slab, k_loc, klb, kub),
                 out=BandedMatrix(C_slab, m_loc, klc, kuc))
        else:
            gbmm_fn(m_loc, n_loc, k_loc, A_slab, kla, kua, B_slab, klb, kub, C_slab, klc, kuc)
        return C_slab


def split_batch(batch: int, world: int, rank: int):
    """Batched independent products (config C5) are replicas: contiguous batch ranges, no communication."""
    per = -(-batch // world)
    b0 = min(batch, rank * per)
    return b0, min(batch, b0 + per)


# ------------------------------------------------------------------------------------------------
# kron(A,I) + kron(I,B) partitioned by grid columns
# ------------------------------------------------------------------------------------------------
class ShardedKronSum:
    """Rank r owns grid columns [r0,r1) of X (n2 x n1, column-major: a contiguous range of
    x = vec(X)); plan = RowShard.plan(n1, kla, kua) over COLUMNS.  It stores A's column slice
    [c0,c1) (a_slab: (c1-c0) x (kla+kua+1)) and all of B; X lives in a ghost-cell buffer of
    n2 x (c1-c0) columns.  Only the ghost columns (n2 values each) are exchanged; B*X is local."""

    def __init__(self, n1, n2, kla, kua, a_slab, B, rank, world):
        self.plan = RowShard.plan(n1, kla, kua, rank, world)
        self.n1, self.n2, self.kla, self.kua = n1, n2, kla, kua
        if tuple(a_slab.shape) != (self.plan.n_loc, kla + kua + 1):
            raise LbmArgumentError(f"A slab must be ({self.plan.n_loc}, {kla + kua + 1})")
        self.a_slab = a_slab.contiguous()
        self.B = B

    def new_buffer(self, dtype, device):
        return torch.zeros((self.plan.n_loc, self.n2), dtype=dtype, device=device)   # column c = row c (n2 contiguous)

    def owned(self, xbuf):
        p = self.plan
        return xbuf[p.left: p.left + p.m_loc]

    def exchange(self, xbuf, group=None):
        p = self.plan
        flat = xbuf.view(-1)
        ops, n2 = [], self.n2
        own0 = p.left * n2
        if p.has_left:
            cnt = min(p.u, p.m_loc) * n2
            if cnt:
                ops.append(dist.P2POp(dist.isend, flat[own0: own0 + cnt], p.rank - 1, group))
            if p.left:
                ops.append(dist.P2POp(dist.irecv, flat[: own0], p.rank - 1, group))
        if p.has_right:
            cnt = min(p.l, p.m_loc) * n2
            end = own0 + p.m_loc * n2
            if cnt:
                ops.append(dist.P2POp(dist.isend, flat[end - cnt: end], p.rank + 1, group))
            if p.right:
                ops.append(dist.P2POp(dist.irecv, flat[end:], p.rank + 1, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def mul_(self, y_own, xbuf, alpha=1.0, beta=0.0, group=None, local_apply=None, overlap=True):
        """y_own ((r1-r0) x n2 tensor = n2 x (r1-r0) column-major) <- alpha*(X A^T + B X)[:, r0:r1] + beta*y_own."""
        p = self.plan
        if local_apply is None and group is None and xbuf.is_cuda:
            from . import _lib
            from .banded import _ft, _pfx, _ptr
            key = (y_own.data_ptr(), xbuf.data_ptr(), float(alpha), float(beta), bool(overlap))
            cached = getattr(self, "_prepared", None)
            if cached is not None and cached[0] == key:      # prepared call (see ShardedBanded.mul_)
                h, fn, args = cached[1]
                h.use_torch_stream()
                rc = fn(*args)
                if rc:
                    h.check(rc, "lbm_kronsum2d_mv_sharded")
                return y_own
            h = _lib.handle_of(xbuf)
            if p.world == 1 or h.comm_world == p.world:
                fn = getattr(_lib.lib(), f"lbm_{_pfx(xbuf.dtype)}kronsum2d_mv_sharded")
                ft = _ft(xbuf.dtype)
                B = self.B
                args = (h.ptr, self.n1, self.n2, self.kla, self.kua, _ptr(self.a_slab), self.kla + self.kua + 1,
                        B.l, B.u, _ptr(B.data), B.lda, ft(alpha), _ptr(xbuf), ft(beta), _ptr(y_own), 1 if overlap else 0)
                self._prepared = (key, (h, fn, args))
                h.check(fn(*args), "lbm_kronsum2d_mv_sharded")
                return y_own
            raise LbmArgumentError("call lbm_b200.shard.init_comm() first (or pass a local_apply for host-logic tests)")
        if p.world > 1:
            self.exchange(xbuf, group)
        return local_apply(y_own, self, xbuf, alpha, beta)


# ------------------------------------------------------------------------------------------------
# Tridiagonal / SymTridiagonal / Bidiagonal mat-vec, row-sharded (BASELINE config 2)
# ------------------------------------------------------------------------------------------------
class ShardedTridiagonal:
    """Rank r owns rows [r0,r1) of an n x n Tridiagonal(dl,d,du) (SymTridiagonal: dl is du;
    Bidiagonal: one of them None).  It keeps the slices of the diagonals over the ghost-extended
    index range [c0,c1) = [r0-1, r1+1) ∩ [0,n) and x in a ghost-cell vector of the same range; after
    the one-element ghost exchange the local apply is the ordinary SQUARE tridiagonal kernel on the
    extended range -- its owned rows are exact, its (at most two) ghost rows are discarded."""

    def __init__(self, n, dl_ext, d_ext, du_ext, rank, world):
        self.plan = RowShard.plan(n, 1 if dl_ext is not None else 0, 1 if du_ext is not None else 0, rank, world)
        p = self.plan
        if d_ext.shape[0] != p.n_loc:
            raise LbmArgumentError(f"d slice must cover [c0,c1) = {p.n_loc} entries")
        for off in (dl_ext, du_ext):
            if off is not None and off.shape[0] != max(p.n_loc - 1, 0):
                raise LbmArgumentError(f"off-diagonal slices must have {max(p.n_loc - 1, 0)} entries")
        self.dl, self.d, self.du = dl_ext, d_ext, du_ext
        self._ybuf = None

    @classmethod
    def slices(cls, n, l, u, rank, world):
        """(c0, c1): global index range of the d / x slices this rank holds; off-diagonal slices are
        dl[c0 : c1-1], du[c0 : c1-1] (dl[i] = A[i+1,i], du[i] = A[i,i+1])."""
        p = RowShard.plan(n, l, u, rank, world)
        return p.c0, p.c1

    def mul_(self, y_own, xbuf, alpha=1.0, beta=0.0, group=None, local_apply=None, overlap=True):
        p = self.plan
        if local_apply is None and group is None and xbuf.is_cuda:
            from . import _lib
            from .banded import _ft, _pfx, _ptr
            key = (y_own.data_ptr(), xbuf.data_ptr(), float(alpha), float(beta), bool(overlap))
            cached = getattr(self, "_prepared", None)
            if cached is not None and cached[0] == key:      # prepared call (see ShardedBanded.mul_)
                h, fn, args = cached[1]
                h.use_torch_stream()
                rc = fn(*args)
                if rc:
                    h.check(rc, "lbm_tridiag_mv_sharded")
                return y_own
            h = _lib.handle_of(xbuf)
            if p.world > 1 and h.comm_world != p.world:
                raise LbmArgumentError("call lbm_b200.shard.init_comm() first")
            # ONE call: the one-element ghost exchange (fused into the launch when the peer mailboxes are up) + the
            # ghost-extended square system restricted to the owned rows, written straight into y_own
            fn = getattr(_lib.lib(), f"lbm_{_pfx(xbuf.dtype)}tridiag_mv_sharded")
            ft = _ft(xbuf.dtype)
            args = (h.ptr, p.n, _ptr(self.dl), _ptr(self.d), _ptr(self.du), ft(alpha), _ptr(xbuf), ft(beta), _ptr(y_own),
                    1 if overlap else 0)
            self._prepared = (key, (h, fn, args))
            h.check(fn(*args), "lbm_tridiag_mv_sharded")
            return y_own
        if self._ybuf is None or self._ybuf.shape[0] != p.n_loc or self._ybuf.device != xbuf.device:
            self._ybuf = torch.empty(p.n_loc, dtype=xbuf.dtype, device=xbuf.device)
        yb = self._ybuf
        if beta != 0.0:
            yb[p.left: p.left + p.m_loc].copy_(y_own)
        if p.world > 1:
            p.exchange(xbuf, group)
        local_apply(yb, p.n_loc, self.dl, self.d, self.du, xbuf, alpha, beta)
        y_own.copy_(yb[p.left: p.left + p.m_loc])
        return y_own


def C_void(t):
    import ctypes as C
    return C.c_void_p(t.data_ptr())
