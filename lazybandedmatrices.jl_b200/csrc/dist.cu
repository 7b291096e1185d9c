// dist.cu -- row-sharded banded mul! across the GPUs of one box, one process per GPU.
//
// The only communication on the path is the nearest-neighbour exchange of the max(kl,ku)-wide
// ghosts of x.  It lives INSIDE the library so that a sharded mul! is one C call:
//   x ready (event) -> [comm stream] ncclGroup{send/recv ghosts} -> event
//                   -> [main stream] interior rows (need no ghosts, overlap the exchange)
//                   -> wait event  -> the <= kl+ku boundary rows.
// Two transports for the ghosts:
//  * peer mailboxes (default once lbm_peer_connect has run): every rank owns a small mailbox block that its two
//    neighbours map through CUDA IPC.  `halo_push_kernel` writes this rank's boundary values straight into the
//    neighbours' mailboxes with NVLink peer stores and then publishes the epoch number in their flag words
//    (system-scope release); `halo_pull_kernel` spins on the local flags (acquire), then moves the mailbox contents
//    into the ghost cells.  Mailbox slots are double-buffered by epoch parity, which -- because both directions of a
//    neighbour pair signal every epoch -- makes an acknowledgement round unnecessary.  Two tiny kernels on the comm
//    stream replace the NCCL group (~20 us -> a few us per apply; the payloads are latency-, not bandwidth-bound).
//  * ncclSend/ncclRecv (fallback: no peer access, halo larger than a mailbox slot, exchange_mode = 1).
// Since round 2 the DEFAULT is neither: the exchange is fused into the compute launch (halo.cuh) -- the streaming kernel
// itself pushes the boundary values at its start and patches the ghost part of its last tiles' x windows in shared
// memory, so a sharded apply is ONE launch on ONE stream.  The two-kernel exchange below remains for shapes the fused
// kernels do not cover (wide bands, op T, unaligned operands), for `overlap = 0`, and for exchange_fuse = 2; it speaks
// the same mailbox protocol, so ranks may mix the two within one epoch.
// NCCL is dlopen'ed (libnccl.so.2: the copy torch already loaded in-process, else the system
// one), so the library has no link-time dependency and builds without NCCL headers.
#include "halo.cuh"

#include <dlfcn.h>

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(void **, int, nccl_uid, int);
typedef int (*fn_destroy)(void *);
typedef int (*fn_send)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_recv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_group)(void);
typedef const char *(*fn_errstr)(int);

struct NcclApi {
    void *lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_send send = nullptr;
    fn_recv recv = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    fn_errstr errstr = nullptr;
    bool ok = false;
} g_nccl;

bool nccl_load() {
    if (g_nccl.ok) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return false;
    g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.destroy = (fn_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.send = (fn_send)dlsym(g_nccl.lib, "ncclSend");
    g_nccl.recv = (fn_recv)dlsym(g_nccl.lib, "ncclRecv");
    g_nccl.group_start = (fn_group)dlsym(g_nccl.lib, "ncclGroupStart");
    g_nccl.group_end = (fn_group)dlsym(g_nccl.lib, "ncclGroupEnd");
    g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
    g_nccl.ok = g_nccl.get_uid && g_nccl.init_rank && g_nccl.destroy && g_nccl.send && g_nccl.recv &&
                g_nccl.group_start && g_nccl.group_end;
    return g_nccl.ok;
}

int fail_nccl(lbm_handle_t h, int rc, const char *what) {
    if (h) snprintf(h->err, 512, "NCCL error %d (%s) in %s", rc, g_nccl.errstr ? g_nccl.errstr(rc) : "?", what);
    return 1000 + rc;
}
#define LBM_NCCL(h, call)                              \
    do {                                               \
        int r__ = (call);                              \
        if (r__ != 0) return fail_nccl(h, r__, #call); \
    } while (0)

struct Plan {
    int64_t r0, r1, c0, c1, left, right, kl_loc, ku_loc, m_loc, n_loc;
    int64_t per;        // rows per full shard
    int world_eff;      // ranks that own rows: ceil(n / per); ranks >= world_eff are empty and have no neighbours
    int has_left, has_right;
};

// contiguous near-equal row blocks whose starts are multiples of 4 (keeps every shard's slab and
// vectors 16-byte aligned for the bulk-TMA kernels); same rule as lbm_b200/shard.py:split_rows.
// Rounding `per` up can leave trailing ranks EMPTY (n = 100, world = 8: rank 6 has four rows, rank 7 none) and the last
// non-empty rank short.  Both are legal: the effective world shrinks to the ranks that own rows -- an empty rank has no
// neighbours and does not take part in the exchange, its left neighbour treats itself as the last rank -- and a short
// LAST shard is fine because the ghosts it serves are clipped at n.  What is NOT representable is a ghost that spans two
// ranks (per < max(kl,ku) with more than one non-empty rank): plan_valid() rejects that on EVERY rank alike, so no rank
// ever returns while a neighbour still waits for it.
Plan make_plan(int64_t n, int64_t kl, int64_t ku, int rank, int world) {
    Plan p;
    int64_t per = (n + world - 1) / world;
    per = (per + 3) / 4 * 4;
    if (per < 4) per = 4;
    p.per = per;
    p.world_eff = (int)((n + per - 1) / per);
    if (p.world_eff < 1) p.world_eff = 1;
    p.r0 = (int64_t)rank * per < n ? (int64_t)rank * per : n;
    p.r1 = p.r0 + per < n ? p.r0 + per : n;
    p.c0 = p.r0 - kl > 0 ? p.r0 - kl : 0;
    p.c1 = p.r1 + ku < n ? p.r1 + ku : n;
    p.left = p.r0 - p.c0;
    p.right = p.c1 - p.r1;
    p.kl_loc = kl - p.left;
    p.ku_loc = ku + p.left;
    p.m_loc = p.r1 - p.r0;
    p.n_loc = p.c1 - p.c0;
    p.has_left = rank > 0 && rank < p.world_eff;
    p.has_right = rank < p.world_eff - 1;
    if (p.m_loc == 0) {   // owns nothing: an empty column slice too
        p.c0 = p.c1 = p.r0;
        p.left = p.right = p.n_loc = 0;
        p.kl_loc = kl;
        p.ku_loc = ku;
    }
    return p;
}

// identical on every rank: a pure function of (n, kl, ku, world)
bool plan_valid(const Plan &p, int64_t kl, int64_t ku) { return p.world_eff <= 1 || p.per >= (kl > ku ? kl : ku); }

// ---- peer-mailbox transport -----------------------------------------------------------------------------------------
// flag words: [0, NF) = epochs published by the LEFT neighbour, [NF, 2NF) = by the RIGHT one.  The 1-D exchanges use word 0
// of each side; the grid-column exchange of the Kronecker-sum kernel uses one word per 256-row strip (kron.cu).
constexpr int PEER_NFLAGS = 4096;
constexpr size_t PEER_FLAG_BYTES = (size_t)2 * PEER_NFLAGS * sizeof(unsigned long long);

inline char *peer_slot(void *block, size_t slot_bytes, int from_side, int parity) {
    return reinterpret_cast<char *>(block) + PEER_FLAG_BYTES + (size_t)(2 * from_side + parity) * slot_bytes;
}

inline unsigned long long halo_timeout_ns(lbm_handle_t h) {
    const long long ms = h->tune.halo_timeout_ms > 0 ? h->tune.halo_timeout_ms : 30000;
    return (unsigned long long)ms * 1000000ull;
}

// Everything one epoch of the mailbox protocol needs, for the fused kernels and for the two-kernel exchange alike.
// Advances the epoch.  `unit` = bytes per ghost unit (an element, or a grid column of n2 elements).
lbm::HaloFuse make_fuse(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku, char *xbuf, int64_t unit, int nflags = 1) {
    lbm::HaloFuse f;
    memset(&f, 0, sizeof(f));
    f.on = 1;
    f.elem_bytes = (int)unit;
    f.nflags = nflags >= 1 && nflags <= PEER_NFLAGS ? nflags : 1;
    const unsigned long long epoch = ++h->peer_epoch;
    const int par = (int)(epoch & 1);
    f.epoch = epoch;
    f.timeout_ns = halo_timeout_ns(h);
    f.status = h->status_dev;
    char *own = xbuf + p.left * unit;
    unsigned long long *lflags = reinterpret_cast<unsigned long long *>(h->peer_local);
    if (p.has_left) {   // left neighbour: gets my first ku entries (its right ghosts), sends me my left ghosts
        const int64_t cnt = ku < p.m_loc ? ku : p.m_loc;
        f.push_src[0] = own;
        f.push_dst[0] = peer_slot(h->peer_remote[0], h->peer_slot_bytes, /*from its RIGHT*/ 1, par);
        f.push_flag[0] = reinterpret_cast<unsigned long long *>(h->peer_remote[0]) + PEER_NFLAGS;
        f.push_bytes[0] = cnt * unit;
        f.pull_src[0] = peer_slot(h->peer_local, h->peer_slot_bytes, 0, par);
        f.pull_flag[0] = lflags + 0;
        f.ghost_begin[0] = 0;
        f.ghost_end[0] = p.left;
        f.wait_only[0] = p.left == 0;
    }
    if (p.has_right) {   // right neighbour: gets my last kl entries, sends me my right ghosts
        const int64_t cnt = kl < p.m_loc ? kl : p.m_loc;
        f.push_src[1] = own + (p.m_loc - cnt) * unit;
        f.push_dst[1] = peer_slot(h->peer_remote[1], h->peer_slot_bytes, /*from its LEFT*/ 0, par);
        f.push_flag[1] = reinterpret_cast<unsigned long long *>(h->peer_remote[1]) + 0;
        f.push_bytes[1] = cnt * unit;
        f.pull_src[1] = peer_slot(h->peer_local, h->peer_slot_bytes, 1, par);
        f.pull_flag[1] = lflags + PEER_NFLAGS;
        f.ghost_begin[1] = p.left + p.m_loc;
        f.ghost_end[1] = p.left + p.m_loc + p.right;
        f.wait_only[1] = p.right == 0;
    }
    return f;
}

// the peer transport can carry this exchange (uniform across ranks: mode, connection state and halo size are)
bool peer_usable(lbm_handle_t h, int64_t kl, int64_t ku, int64_t unit) {
    const int64_t worst = (kl > ku ? kl : ku) * unit;
    return h->peer_on && h->tune.exchange_mode != 1 && (size_t)worst <= h->peer_slot_bytes;
}

// ---- the two-kernel exchange (unfused) -------------------------------------------------------------------------------
// grid = 1 CTA.  Payload first (peer stores), then fence + barrier, then the flags.
__global__ void __launch_bounds__(256) halo_push_kernel(const lbm::HaloFuse f) { lbm::halo_push(f, threadIdx.x, 256); }

// grid = 2 CTAs (one per side).  Wait for the neighbour's epoch (wall-clock bounded: a late neighbour is not a fault, a lost
// one sets the handle's fault word instead of trapping), then mailbox -> ghost cells (read around L1).
__global__ void __launch_bounds__(256) halo_pull_kernel(const lbm::HaloFuse f, char *xbuf) {
    const int s = blockIdx.x;
    if (!f.pull_flag[s]) return;
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < f.nflags; i += blockDim.x)   // a strip-wise pusher publishes one word per strip
        if (!lbm::halo_wait_flag(f.pull_flag[s] + i, f.epoch, f.timeout_ns, f.status)) ok = 0;
    __syncthreads();
    const int64_t nb = (f.ghost_end[s] - f.ghost_begin[s]) * f.elem_bytes;
    if (!ok || nb <= 0) return;
    const char *src = f.pull_src[s];
    char *dst = xbuf + f.ghost_begin[s] * f.elem_bytes;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)nb) & 15u) == 0) {
        const int4 *s4 = reinterpret_cast<const int4 *>(src);
        int4 *d4 = reinterpret_cast<int4 *>(dst);
        for (int64_t i = threadIdx.x; i < nb / 16; i += blockDim.x) d4[i] = __ldcg(s4 + i);
    } else {
        for (int64_t i = threadIdx.x; i < nb; i += blockDim.x) dst[i] = __ldcg(src + i);
    }
}

// same contract as exchange_nccl(): ghosts of xbuf valid once h->ev_halo has fired
int exchange_peer(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku, char *xbuf, int64_t eb, int nflags) {
    LBM_CUDA(h, cudaEventRecord(h->ev_ready, h->stream));
    LBM_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
    const lbm::HaloFuse f = make_fuse(h, p, kl, ku, xbuf, eb, nflags);
    halo_push_kernel<<<1, 256, 0, h->comm_stream>>>(f);
    LBM_LAUNCH_CHECK(h, "halo_push_kernel");
    halo_pull_kernel<<<2, 256, 0, h->comm_stream>>>(f, xbuf);
    LBM_LAUNCH_CHECK(h, "halo_pull_kernel");
    LBM_CUDA(h, cudaEventRecord(h->ev_halo, h->comm_stream));
    return 0;
}

int exchange_nccl(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku, char *xbuf, int64_t eb) {
    if (!h->nccl_comm) return lbm_fail_arg(h, 1, "no NCCL communicator on this handle (and the peer transport cannot carry this halo)");
    // sends read / recvs write xbuf: order them after everything already queued on the main stream
    LBM_CUDA(h, cudaEventRecord(h->ev_ready, h->stream));
    LBM_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
    char *own = xbuf + p.left * eb;
    LBM_NCCL(h, g_nccl.group_start());
    if (p.has_left) {
        const int64_t cnt = ku < p.m_loc ? ku : p.m_loc;  // previous rank's right ghost = my first ku entries
        if (cnt > 0) LBM_NCCL(h, g_nccl.send(own, (size_t)(cnt * eb), 0 /*ncclInt8*/, h->rank - 1, h->nccl_comm, h->comm_stream));
        if (p.left > 0) LBM_NCCL(h, g_nccl.recv(xbuf, (size_t)(p.left * eb), 0, h->rank - 1, h->nccl_comm, h->comm_stream));
    }
    if (p.has_right) {
        const int64_t cnt = kl < p.m_loc ? kl : p.m_loc;  // next rank's left ghost = my last kl entries
        if (cnt > 0)
            LBM_NCCL(h, g_nccl.send(own + (p.m_loc - cnt) * eb, (size_t)(cnt * eb), 0, h->rank + 1, h->nccl_comm, h->comm_stream));
        if (p.right > 0)
            LBM_NCCL(h, g_nccl.recv(own + p.m_loc * eb, (size_t)(p.right * eb), 0, h->rank + 1, h->nccl_comm, h->comm_stream));
    }
    LBM_NCCL(h, g_nccl.group_end());
    LBM_CUDA(h, cudaEventRecord(h->ev_halo, h->comm_stream));
    return 0;
}

// A rank beyond the effective world owns nothing and takes no part in an exchange -- but the epoch counter is the
// protocol's clock and must tick on EVERY rank for every exchange the others perform, or a rank that is empty for one
// operator and owns rows of the next (a different n) would rejoin its neighbours out of step.
void skip_exchange(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku, int64_t unit) {
    if (h->world > 1 && p.world_eff > 1 && peer_usable(h, kl, ku, unit)) ++h->peer_epoch;
}

// transport choice; the halo each way is at most max(kl,ku) units.  Only called on ranks that own rows.
int exchange(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku, char *xbuf, int64_t eb, int nflags = 1) {
    if (peer_usable(h, kl, ku, eb)) return exchange_peer(h, p, kl, ku, xbuf, eb, nflags);
    return exchange_nccl(h, p, kl, ku, xbuf, eb);
}

// Common entry checks of every sharded call: uniform argument validation first (identical on all ranks), then the
// device-side fault word.  Returns 0 to proceed.
int sharded_enter(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku) {
    if (h->world > 1 && !h->has_comm) return lbm_fail_arg(h, 1, "lbm_comm_init has not been called on this handle");
    if (!plan_valid(p, kl, ku))
        return lbm_fail_arg(h, 2, "shards narrower than the halo max(kl,ku): use fewer ranks (same verdict on every rank)");
    return lbm_check_fault(h);
}

template <typename T> int gbmv_call(lbm_handle_t, int64_t, int64_t, int64_t, int64_t, T, const T *, int64_t, const T *, T, T *);
template <> int gbmv_call<double>(lbm_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, double alpha, const double *A,
                                  int64_t lda, const double *x, double beta, double *y) {
    return lbm_dgbmv(h, 'N', m, n, kl, ku, alpha, A, lda, x, beta, y);
}
template <> int gbmv_call<float>(lbm_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, float alpha, const float *A,
                                 int64_t lda, const float *x, float beta, float *y) {
    return lbm_sgbmv(h, 'N', m, n, kl, ku, alpha, A, lda, x, beta, y);
}
template <typename T> int gbmv_fused_call(lbm_handle_t, int64_t, int64_t, int64_t, int64_t, T, const T *, int64_t, const T *, T, T *, const lbm::HaloFuse *);
template <> int gbmv_fused_call<double>(lbm_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, double alpha, const double *A,
                                        int64_t lda, const double *x, double beta, double *y, const lbm::HaloFuse *hf) {
    return lbm_internal_dgbmv_fused(h, m, n, kl, ku, alpha, A, lda, x, beta, y, hf);
}
template <> int gbmv_fused_call<float>(lbm_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, float alpha, const float *A,
                                       int64_t lda, const float *x, float beta, float *y, const lbm::HaloFuse *hf) {
    return lbm_internal_sgbmv_fused(h, m, n, kl, ku, alpha, A, lda, x, beta, y, hf);
}
template <typename T> int tri_fused_call(lbm_handle_t, int64_t, const T *, const T *, const T *, T, const T *, T, T *, int64_t, int64_t, const lbm::HaloFuse *);
template <> int tri_fused_call<double>(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du, double alpha,
                                       const double *x, double beta, double *y, int64_t a, int64_t b, const lbm::HaloFuse *hf) {
    return lbm_internal_dtridiag_rows_fused(h, n, dl, d, du, alpha, x, beta, y, a, b, hf);
}
template <> int tri_fused_call<float>(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du, float alpha,
                                      const float *x, float beta, float *y, int64_t a, int64_t b, const lbm::HaloFuse *hf) {
    return lbm_internal_stridiag_rows_fused(h, n, dl, d, du, alpha, x, beta, y, a, b, hf);
}

// rows [a,b) of the local slab: again a rectangular banded matrix (column slice, shifted bandwidths)
template <typename T>
int apply_rows(lbm_handle_t h, const Plan &p, int64_t a, int64_t b, T alpha, const T *A, int64_t lda, const T *xbuf,
               T beta, T *y) {
    if (b <= a) return 0;
    const int64_t c0 = a - p.kl_loc > 0 ? a - p.kl_loc : 0;
    const int64_t c1 = b + p.ku_loc < p.n_loc ? b + p.ku_loc : p.n_loc;
    const int64_t shift = a - c0;
    return gbmv_call<T>(h, b - a, c1 - c0, p.kl_loc - shift, p.ku_loc + shift, alpha, A + c0 * lda, lda, xbuf + c0, beta,
                        y + a);
}

// The fused-exchange launch is wanted (peer transport up, not switched off) -- whether the KERNEL can take it is decided
// by the kernel's own launcher (LBM_NOT_FUSED).
bool want_fused(lbm_handle_t h, int64_t kl, int64_t ku, int64_t unit, int overlap) {
    return overlap && h->tune.exchange_fuse != 2 && peer_usable(h, kl, ku, unit);
}

template <typename T>
int gbmv_sharded_plan(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku, T alpha, const T *A, int64_t lda, T *xbuf, T beta,
                      T *y, int overlap) {
    int rc;
    if (p.world_eff <= 1 || h->world == 1) {
        if (h->tune.exchange_fuse == 3) {   // self-test knob: the HALO instantiation with no neighbours (what the variant itself costs)
            lbm::HaloFuse none;
            memset(&none, 0, sizeof(none));
            none.on = 1 | (int)(h->tune.halo_timeout_ms & 6);   // self-test sub-knobs ride in halo_timeout_ms: 2 = no rotation, 4 = no patch test
            none.nflags = 1;
            rc = gbmv_fused_call<T>(h, p.m_loc, p.n_loc, p.kl_loc, p.ku_loc, alpha, A, lda, xbuf, beta, y, &none);
            if (rc != LBM_NOT_FUSED) return rc;
        }
        return apply_rows<T>(h, p, 0, p.m_loc, alpha, A, lda, xbuf, beta, y);
    }
    if (want_fused(h, kl, ku, sizeof(T), overlap)) {
        const lbm::HaloFuse f = make_fuse(h, p, kl, ku, reinterpret_cast<char *>(xbuf), sizeof(T));
        rc = gbmv_fused_call<T>(h, p.m_loc, p.n_loc, p.kl_loc, p.ku_loc, alpha, A, lda, xbuf, beta, y, &f);
        if (rc != LBM_NOT_FUSED) return rc;
        --h->peer_epoch;   // nothing was launched: this epoch goes through the two-kernel exchange below
    }
    if ((rc = exchange(h, p, kl, ku, reinterpret_cast<char *>(xbuf), (int)sizeof(T)))) return rc;
    if (!overlap) {
        LBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
        return apply_rows<T>(h, p, 0, p.m_loc, alpha, A, lda, xbuf, beta, y);
    }
    int64_t a = p.left == 0 ? 0 : (kl + 3) / 4 * 4;  // aligned start keeps the interior on the bulk-TMA kernel
    int64_t b = p.right == 0 ? p.m_loc : p.m_loc - ku;
    if (a > p.m_loc) a = p.m_loc;
    if (b < a) b = a;
    if ((rc = apply_rows<T>(h, p, a, b, alpha, A, lda, xbuf, beta, y))) return rc;  // overlaps the exchange
    LBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
    if ((rc = apply_rows<T>(h, p, 0, a, alpha, A, lda, xbuf, beta, y))) return rc;
    return apply_rows<T>(h, p, b, p.m_loc, alpha, A, lda, xbuf, beta, y);
}

template <typename T>
int gbmv_sharded(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, T alpha, const T *A, int64_t lda, T *xbuf, T beta,
                 T *y, int overlap) {
    if (!h) return -1;
    if (n < 0) return lbm_fail_arg(h, 2, "n < 0");
    if (kl < 0) return lbm_fail_arg(h, 3, "kl < 0");
    if (ku < 0) return lbm_fail_arg(h, 4, "ku < 0");
    if (lda < kl + ku + 1) return lbm_fail_arg(h, 7, "lda < kl+ku+1");
    lbm_device_guard g(h->device);
    const Plan p = make_plan(n, kl, ku, h->rank, h->world);
    int rc = sharded_enter(h, p, kl, ku);
    if (rc) return rc;
    if (p.m_loc == 0) {   // beyond the effective world: owns nothing, has no neighbours
        skip_exchange(h, p, kl, ku, sizeof(T));
        return 0;
    }
    return gbmv_sharded_plan<T>(h, p, kl, ku, alpha, A, lda, xbuf, beta, y, overlap);
}

// mul!(y, ApplyMatrix(*, F_0, ..., F_{nf-1}), x) row-sharded, ONE call: right to left; the output of stage f lands in the
// owned block of a ghost-cell vector laid out for the factor that consumes it (bandwidths of F_{f-1}), whose apply then
// exchanges that intermediate's halo -- "halos of the intermediate in product chains" (north_star; the flattening rule is
// the one the reference relies on at src/blockkron.jl:208-214).  Temporaries ping-pong in the handle's scratch.
template <typename T>
int gbmv_chain_sharded(lbm_handle_t h, int nf, int64_t n, const int64_t *kl, const int64_t *ku, const T *const *A,
                       const int64_t *lda, T alpha, T *xbuf, T beta, T *y) {
    if (!h) return -1;
    if (nf < 1) return lbm_fail_arg(h, 2, "nf < 1");
    if (n < 0) return lbm_fail_arg(h, 3, "n < 0");
    if (!kl || !ku || !A || !lda) return lbm_fail_arg(h, 4, "NULL factor description");
    for (int f = 0; f < nf; ++f) {
        if (kl[f] < 0 || ku[f] < 0) return lbm_fail_arg(h, 4, "negative bandwidth");
        if (lda[f] < kl[f] + ku[f] + 1) return lbm_fail_arg(h, 7, "lda < kl+ku+1");
    }
    lbm_device_guard g(h->device);
    int rc;
    int64_t pitch = 0;
    for (int f = 0; f < nf; ++f) {   // every factor's plan is validated before anything is queued (uniform verdict)
        const Plan p = make_plan(n, kl[f], ku[f], h->rank, h->world);
        if ((rc = sharded_enter(h, p, kl[f], ku[f]))) return rc;
        if (p.n_loc > pitch) pitch = p.n_loc;
    }
    const Plan p0 = make_plan(n, kl[0], ku[0], h->rank, h->world);
    if (p0.m_loc == 0) {   // one exchange per factor on the ranks that own rows
        for (int f = 0; f < nf; ++f) skip_exchange(h, make_plan(n, kl[f], ku[f], h->rank, h->world), kl[f], ku[f], sizeof(T));
        return 0;
    }
    if (nf == 1) return gbmv_sharded_plan<T>(h, p0, kl[0], ku[0], alpha, A[0], lda[0], xbuf, beta, y, 1);
    pitch = (pitch + 31) / 32 * 32;
    void *scr;
    if ((rc = lbm_scratch(h, (size_t)(2 * pitch) * sizeof(T), &scr))) return rc;
    T *tmp[2] = {(T *)scr, (T *)scr + pitch};
    T *v = xbuf;
    for (int f = nf - 1; f >= 1; --f) {
        const Plan pf = make_plan(n, kl[f], ku[f], h->rank, h->world);
        const Plan pc = make_plan(n, kl[f - 1], ku[f - 1], h->rank, h->world);   // the consumer's ghost layout
        T *t = tmp[f & 1];
        if ((rc = gbmv_sharded_plan<T>(h, pf, kl[f], ku[f], (T)1, A[f], lda[f], v, (T)0, t + pc.left, 1))) return rc;
        v = t;
    }
    return gbmv_sharded_plan<T>(h, p0, kl[0], ku[0], alpha, A[0], lda[0], v, beta, y, 1);
}

// Tridiagonal / SymTridiagonal / Bidiagonal mat-vec, row-sharded, ONE call: the rank holds the diagonals over its
// ghost-extended index range [c0,c1) and applies that square system restricted to the owned rows.
template <typename T>
int tridiag_sharded(lbm_handle_t h, int64_t n, const T *dl, const T *d, const T *du, T alpha, T *xbuf, T beta, T *y, int overlap) {
    if (!h) return -1;
    if (n < 0) return lbm_fail_arg(h, 2, "n < 0");
    lbm_device_guard g(h->device);
    const int64_t kl = dl ? 1 : 0, ku = du ? 1 : 0;
    const Plan p = make_plan(n, kl, ku, h->rank, h->world);
    int rc = sharded_enter(h, p, kl, ku);
    if (rc) return rc;
    if (p.m_loc == 0) {
        skip_exchange(h, p, kl, ku, sizeof(T));
        return 0;
    }
    if (!d) return lbm_fail_arg(h, 4, "d == NULL");
    if (!xbuf) return lbm_fail_arg(h, 7, "xbuf == NULL");
    if (!y) return lbm_fail_arg(h, 9, "y_own == NULL");
    if (p.world_eff > 1 && h->world > 1) {
        if (want_fused(h, kl, ku, sizeof(T), overlap)) {
            const lbm::HaloFuse f = make_fuse(h, p, kl, ku, reinterpret_cast<char *>(xbuf), sizeof(T));
            rc = tri_fused_call<T>(h, p.n_loc, dl, d, du, alpha, xbuf, beta, y, p.left, p.left + p.m_loc, &f);
            if (rc != LBM_NOT_FUSED) return rc;
            --h->peer_epoch;
        }
        if ((rc = exchange(h, p, kl, ku, reinterpret_cast<char *>(xbuf), (int)sizeof(T)))) return rc;
        LBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
    }
    return tri_fused_call<T>(h, p.n_loc, dl, d, du, alpha, xbuf, beta, y, p.left, p.left + p.m_loc, nullptr);
}

}  // namespace

// internal (kron.cu): plan over `n` units of `unit_bytes` each; then either a fused-exchange descriptor for the kernel
// (returns 1, *hf filled, epoch advanced) or the two-kernel / NCCL exchange started on the comm stream (returns 0; the
// main stream does NOT wait -- the caller overlaps interior work and then waits on h->ev_halo).  < 0 / > 1: error code.
// plan10 = {r0, r1, c0, c1, left, right, kl_loc, ku_loc, world_eff, m_loc}.
int lbm_internal_exchange(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, void *buf, int64_t unit_bytes, int64_t *plan10,
                          int overlap, lbm::HaloFuse *hf, int *fused, int nflags) {
    const Plan p = make_plan(n, kl, ku, h->rank, h->world);
    plan10[0] = p.r0; plan10[1] = p.r1; plan10[2] = p.c0; plan10[3] = p.c1;
    plan10[4] = p.left; plan10[5] = p.right; plan10[6] = p.kl_loc; plan10[7] = p.ku_loc;
    plan10[8] = p.world_eff; plan10[9] = p.m_loc;
    *fused = 0;
    int rc = sharded_enter(h, p, kl, ku);
    if (rc) return rc;
    if (p.m_loc == 0) skip_exchange(h, p, kl, ku, unit_bytes);
    if (h->world == 1 || p.world_eff <= 1 || p.m_loc == 0) return 0;
    if (hf && want_fused(h, kl, ku, unit_bytes, overlap)) {
        *hf = make_fuse(h, p, kl, ku, reinterpret_cast<char *>(buf), unit_bytes, nflags);
        *fused = 1;
        return 0;
    }
    return exchange(h, p, kl, ku, reinterpret_cast<char *>(buf), unit_bytes, nflags);
}
// comm stream + events of a handle that takes part in a row partition (created once)
int lbm_internal_comm_resources(lbm_handle_t h) {
    if (h->comm_stream) return 0;
    LBM_CUDA(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    LBM_CUDA(h, cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    LBM_CUDA(h, cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
    return 0;
}
// this rank's mailbox block: [flags 256 B][from-left slots 0,1][from-right slots 0,1]
int lbm_internal_peer_alloc_local(lbm_handle_t h, size_t slot_bytes) {
    const size_t total = PEER_FLAG_BYTES + 4 * slot_bytes;
    LBM_CUDA(h, cudaMalloc(&h->peer_local, total));
    LBM_CUDA(h, cudaMemset(h->peer_local, 0, total));
    LBM_CUDA(h, cudaDeviceSynchronize());
    h->peer_slot_bytes = slot_bytes;
    h->peer_epoch = 0;
    return 0;
}
// the kernel turned the fused launch down (LBM_NOT_FUSED): give the epoch back and run the unfused exchange
int lbm_internal_exchange_unfuse(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, void *buf, int64_t unit_bytes, int nflags) {
    --h->peer_epoch;
    const Plan p = make_plan(n, kl, ku, h->rank, h->world);
    return exchange(h, p, kl, ku, reinterpret_cast<char *>(buf), unit_bytes, nflags);
}

extern "C" {

int lbm_shard_plan(int64_t n, int64_t kl, int64_t ku, int rank, int world, int64_t *out8) {
    if (n < 0) return -1;
    if (kl < 0) return -2;
    if (ku < 0) return -3;
    if (world < 1) return -5;
    if (rank < 0 || rank >= world) return -4;
    if (!out8) return -6;
    const Plan p = make_plan(n, kl, ku, rank, world);
    if (!plan_valid(p, kl, ku)) return -2;   // shards narrower than the halo: the same verdict for every rank
    out8[0] = p.r0; out8[1] = p.r1; out8[2] = p.c0; out8[3] = p.c1;
    out8[4] = p.left; out8[5] = p.right; out8[6] = p.kl_loc; out8[7] = p.ku_loc;
    return 0;
}

int lbm_comm_unique_id(void *id128) {
    if (!id128) return -1;
    if (!nccl_load()) return 1;
    nccl_uid id;
    int rc = g_nccl.get_uid(&id);
    if (rc) return 1000 + rc;
    memcpy(id128, id.internal, 128);
    return 0;
}

int lbm_comm_init(lbm_handle_t h, const void *id128, int rank, int world) {
    if (!h) return -1;
    if (!id128) return lbm_fail_arg(h, 2, "id128 == NULL");
    if (world < 1) return lbm_fail_arg(h, 4, "world < 1");
    if (rank < 0 || rank >= world) return lbm_fail_arg(h, 3, "rank out of range");
    if (!nccl_load()) {
        snprintf(h->err, 512, "could not dlopen libnccl.so.2");
        return 1;
    }
    lbm_device_guard g(h->device);
    lbm_comm_destroy(h);
    nccl_uid id;
    memcpy(id.internal, id128, 128);
    void *comm = nullptr;
    LBM_NCCL(h, g_nccl.init_rank(&comm, world, id, rank));
    h->nccl_comm = comm;
    h->rank = rank;
    h->world = world;
    h->has_comm = 1;
    return lbm_internal_comm_resources(h);
}

// ---- peer-memory transport setup (all three are collective in spirit: every rank of the row partition calls them) ----
int lbm_peer_alloc(lbm_handle_t h, size_t slot_bytes, void *ipc_handle64) {
    if (!h) return -1;
    if (!ipc_handle64) return lbm_fail_arg(h, 3, "ipc_handle64 == NULL");
    if (slot_bytes == 0 || slot_bytes % 16) return lbm_fail_arg(h, 2, "slot_bytes must be a positive multiple of 16");
    lbm_device_guard g(h->device);
    lbm_peer_disconnect(h);
    int rc = lbm_internal_peer_alloc_local(h, slot_bytes);
    if (rc) return rc;
    cudaIpcMemHandle_t ih;
    LBM_CUDA(h, cudaIpcGetMemHandle(&ih, h->peer_local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    memcpy(ipc_handle64, &ih, 64);
    return 0;
}

int lbm_peer_connect(lbm_handle_t h, const void *left_handle64, const void *right_handle64) {
    if (!h) return -1;
    if (!h->peer_local) return lbm_fail_arg(h, 1, "lbm_peer_alloc has not been called on this handle");
    if (!h->has_comm) return lbm_fail_arg(h, 1, "lbm_comm_init has not been called on this handle");
    if ((h->rank > 0) != (left_handle64 != nullptr)) return lbm_fail_arg(h, 2, "left handle must be given iff rank > 0");
    if ((h->rank < h->world - 1) != (right_handle64 != nullptr)) return lbm_fail_arg(h, 3, "right handle must be given iff rank < world-1");
    lbm_device_guard g(h->device);
    const void *hs[2] = {left_handle64, right_handle64};
    for (int s = 0; s < 2; ++s) {
        if (!hs[s]) continue;
        cudaIpcMemHandle_t ih;
        memcpy(&ih, hs[s], 64);
        LBM_CUDA(h, cudaIpcOpenMemHandle(&h->peer_remote[s], ih, cudaIpcMemLazyEnablePeerAccess));
    }
    h->peer_on = 1;
    h->peer_direct = 0;
    return 0;
}

int lbm_peer_disconnect(lbm_handle_t h) {
    if (!h) return -1;
    lbm_device_guard g(h->device);
    if (h->peer_on || h->peer_local) cudaDeviceSynchronize();
    for (int s = 0; s < 2; ++s) {
        if (h->peer_remote[s] && !h->peer_direct) cudaIpcCloseMemHandle(h->peer_remote[s]);
        h->peer_remote[s] = nullptr;
    }
    if (h->peer_local) cudaFree(h->peer_local);
    h->peer_local = nullptr;
    h->peer_on = 0;
    h->peer_direct = 0;
    h->peer_slot_bytes = 0;
    return 0;
}

int lbm_peer_active(lbm_handle_t h) { return h && h->peer_on && h->tune.exchange_mode != 1 ? 1 : 0; }

int lbm_comm_destroy(lbm_handle_t h) {
    if (!h) return -1;
    lbm_peer_disconnect(h);
    lbm_device_guard g(h->device);
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    if (h->nccl_comm) {
        g_nccl.destroy(h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    if (h->comm_stream) {
        cudaEventDestroy(h->ev_ready);
        cudaEventDestroy(h->ev_halo);
        cudaStreamDestroy(h->comm_stream);
        h->comm_stream = nullptr;
        h->ev_ready = h->ev_halo = nullptr;
    }
    h->rank = 0;
    h->world = 1;
    h->has_comm = 0;
    return 0;
}

int lbm_halo_exchange(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, void *xbuf, int elem_bytes) {
    if (!h) return -1;
    if (n < 0) return lbm_fail_arg(h, 2, "n < 0");
    if (kl < 0) return lbm_fail_arg(h, 3, "kl < 0");
    if (ku < 0) return lbm_fail_arg(h, 4, "ku < 0");
    if (elem_bytes != 4 && elem_bytes != 8) return lbm_fail_arg(h, 6, "elem_bytes must be 4 or 8");
    if (h->world == 1) return 0;
    lbm_device_guard g(h->device);
    const Plan p = make_plan(n, kl, ku, h->rank, h->world);
    int rc = sharded_enter(h, p, kl, ku);   // same verdict on every rank
    if (rc) return rc;
    if (p.m_loc == 0) skip_exchange(h, p, kl, ku, elem_bytes);
    if (p.m_loc == 0 || p.world_eff <= 1) return 0;   // owns nothing / is alone: no neighbours, nothing to exchange
    if ((rc = exchange(h, p, kl, ku, reinterpret_cast<char *>(xbuf), elem_bytes))) return rc;
    LBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
    return 0;
}

int lbm_dgbmv_sharded(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, double alpha, const double *A_slab, int64_t lda,
                      double *xbuf, double beta, double *y_own, int overlap) {
    return gbmv_sharded<double>(h, n, kl, ku, alpha, A_slab, lda, xbuf, beta, y_own, overlap);
}
int lbm_sgbmv_sharded(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, float alpha, const float *A_slab, int64_t lda,
                      float *xbuf, float beta, float *y_own, int overlap) {
    return gbmv_sharded<float>(h, n, kl, ku, alpha, A_slab, lda, xbuf, beta, y_own, overlap);
}
int lbm_dgbmv_chain_sharded(lbm_handle_t h, int nf, int64_t n, const int64_t *kl, const int64_t *ku, const double *const *A_slab,
                            const int64_t *lda, double alpha, double *xbuf, double beta, double *y_own) {
    return gbmv_chain_sharded<double>(h, nf, n, kl, ku, A_slab, lda, alpha, xbuf, beta, y_own);
}
int lbm_sgbmv_chain_sharded(lbm_handle_t h, int nf, int64_t n, const int64_t *kl, const int64_t *ku, const float *const *A_slab,
                            const int64_t *lda, float alpha, float *xbuf, float beta, float *y_own) {
    return gbmv_chain_sharded<float>(h, nf, n, kl, ku, A_slab, lda, alpha, xbuf, beta, y_own);
}
int lbm_dtridiag_mv_sharded(lbm_handle_t h, int64_t n, const double *dl_ext, const double *d_ext, const double *du_ext, double alpha,
                            double *xbuf, double beta, double *y_own, int overlap) {
    return tridiag_sharded<double>(h, n, dl_ext, d_ext, du_ext, alpha, xbuf, beta, y_own, overlap);
}
int lbm_stridiag_mv_sharded(lbm_handle_t h, int64_t n, const float *dl_ext, const float *d_ext, const float *du_ext, float alpha,
                            float *xbuf, float beta, float *y_own, int overlap) {
    return tridiag_sharded<float>(h, n, dl_ext, d_ext, du_ext, alpha, xbuf, beta, y_own, overlap);
}
}
