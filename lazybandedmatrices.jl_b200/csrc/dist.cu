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
// NCCL is dlopen'ed (libnccl.so.2: the copy torch already loaded in-process, else the system
// one), so the library has no link-time dependency and builds without NCCL headers.
#include "lbm_common.cuh"

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
};

// contiguous near-equal row blocks whose starts are multiples of 4 (keeps every shard's slab and
// vectors 16-byte aligned for the bulk-TMA kernels); same rule as lbm_b200/shard.py:split_rows
Plan make_plan(int64_t n, int64_t kl, int64_t ku, int rank, int world) {
    Plan p;
    int64_t per = (n + world - 1) / world;
    per = (per + 3) / 4 * 4;
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
    return p;
}

// ---- peer-mailbox transport -----------------------------------------------------------------------------------------
constexpr size_t PEER_FLAG_BYTES = 256;   // flags[0] = epoch published by the LEFT neighbour, flags[1] = by the RIGHT one

struct PushArgs {
    // side 0: to the left neighbour (it receives into its from-RIGHT mailbox), side 1: to the right neighbour
    const char *src[2];
    char *dst[2];                 // peer mailbox slot (NULL: no such neighbour)
    unsigned long long *flag[2];  // peer flag word to publish the epoch in
    int64_t bytes[2];
    unsigned long long epoch;
};
struct PullArgs {
    const char *src[2];           // local mailbox slot written by the left / right neighbour
    char *dst[2];                 // ghost cells (NULL: nothing to receive on that side)
    const unsigned long long *flag[2];
    int64_t bytes[2];
    int wait[2];                  // a neighbour exists on that side
    unsigned long long epoch;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// grid = 2 CTAs (one per side).  Payload first (peer stores), then fence + barrier, then the flag.
__global__ void __launch_bounds__(256) halo_push_kernel(const PushArgs a) {
    const int s = blockIdx.x;
    if (!a.flag[s]) return;
    const int64_t nb = a.bytes[s];
    if (nb > 0) {
        const char *src = a.src[s];
        char *dst = a.dst[s];
        if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)nb) & 15u) == 0) {
            const int4 *s4 = reinterpret_cast<const int4 *>(src);
            int4 *d4 = reinterpret_cast<int4 *>(dst);
            for (int64_t i = threadIdx.x; i < nb / 16; i += blockDim.x) d4[i] = s4[i];
        } else {
            for (int64_t i = threadIdx.x; i < nb; i += blockDim.x) dst[i] = src[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(a.flag[s], a.epoch);
}

// grid = 2 CTAs.  Wait for the neighbour's epoch, then mailbox -> ghost cells (read around L1).
__global__ void __launch_bounds__(256) halo_pull_kernel(const PullArgs a) {
    const int s = blockIdx.x;
    if (!a.wait[s]) return;
    if (threadIdx.x == 0) {
        unsigned long long spins = 0;
        while (ld_acquire_sys(a.flag[s]) < a.epoch) {
            // a lost neighbour must surface as a CUDA error, never as a hung GPU (~10+ s of polling)
            if (++spins > (1ull << 25)) __trap();
            __nanosleep(64);
        }
    }
    __syncthreads();
    const int64_t nb = a.bytes[s];
    if (nb <= 0 || !a.dst[s]) return;
    const char *src = a.src[s];
    char *dst = a.dst[s];
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)nb) & 15u) == 0) {
        const int4 *s4 = reinterpret_cast<const int4 *>(src);
        int4 *d4 = reinterpret_cast<int4 *>(dst);
        for (int64_t i = threadIdx.x; i < nb / 16; i += blockDim.x) d4[i] = __ldcg(s4 + i);
    } else {
        for (int64_t i = threadIdx.x; i < nb; i += blockDim.x) dst[i] = __ldcg(src + i);
    }
}

inline char *peer_slot(void *block, size_t slot_bytes, int from_side, int parity) {
    return reinterpret_cast<char *>(block) + PEER_FLAG_BYTES + (size_t)(2 * from_side + parity) * slot_bytes;
}

// same contract as exchange(): ghosts of xbuf valid once h->ev_halo has fired
int exchange_peer(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku, char *xbuf, int64_t eb) {
    LBM_CUDA(h, cudaEventRecord(h->ev_ready, h->stream));
    LBM_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
    const unsigned long long epoch = ++h->peer_epoch;
    const int par = (int)(epoch & 1);
    char *own = xbuf + p.left * eb;
    PushArgs pa;
    PullArgs pl;
    memset(&pa, 0, sizeof(pa));
    memset(&pl, 0, sizeof(pl));
    pa.epoch = pl.epoch = epoch;
    unsigned long long *lflags = reinterpret_cast<unsigned long long *>(h->peer_local);
    if (h->rank > 0) {   // left neighbour: gets my first ku entries (its right ghosts), sends me my left ghosts
        const int64_t cnt = ku < p.m_loc ? ku : p.m_loc;
        pa.src[0] = own;
        pa.dst[0] = peer_slot(h->peer_remote[0], h->peer_slot_bytes, /*from its RIGHT*/ 1, par);
        pa.flag[0] = reinterpret_cast<unsigned long long *>(h->peer_remote[0]) + 1;
        pa.bytes[0] = cnt * eb;
        pl.wait[0] = 1;
        pl.flag[0] = lflags + 0;
        pl.src[0] = peer_slot(h->peer_local, h->peer_slot_bytes, 0, par);
        pl.dst[0] = p.left > 0 ? xbuf : nullptr;
        pl.bytes[0] = p.left * eb;
    }
    if (h->rank < h->world - 1) {   // right neighbour: gets my last kl entries, sends me my right ghosts
        const int64_t cnt = kl < p.m_loc ? kl : p.m_loc;
        pa.src[1] = own + (p.m_loc - cnt) * eb;
        pa.dst[1] = peer_slot(h->peer_remote[1], h->peer_slot_bytes, /*from its LEFT*/ 0, par);
        pa.flag[1] = reinterpret_cast<unsigned long long *>(h->peer_remote[1]) + 0;
        pa.bytes[1] = cnt * eb;
        pl.wait[1] = 1;
        pl.flag[1] = lflags + 1;
        pl.src[1] = peer_slot(h->peer_local, h->peer_slot_bytes, 1, par);
        pl.dst[1] = p.right > 0 ? own + p.m_loc * eb : nullptr;
        pl.bytes[1] = p.right * eb;
    }
    halo_push_kernel<<<2, 256, 0, h->comm_stream>>>(pa);
    LBM_LAUNCH_CHECK(h, "halo_push_kernel");
    halo_pull_kernel<<<2, 256, 0, h->comm_stream>>>(pl);
    LBM_LAUNCH_CHECK(h, "halo_pull_kernel");
    LBM_CUDA(h, cudaEventRecord(h->ev_halo, h->comm_stream));
    return 0;
}

int exchange_nccl(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku, char *xbuf, int64_t eb);

// transport choice; the halo each way is at most max(kl,ku) units
int exchange(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku, char *xbuf, int64_t eb) {
    const int64_t worst = (kl > ku ? kl : ku) * eb;
    if (h->peer_on && h->tune.exchange_mode != 1 && (size_t)worst <= h->peer_slot_bytes)
        return exchange_peer(h, p, kl, ku, xbuf, eb);
    return exchange_nccl(h, p, kl, ku, xbuf, eb);
}

int exchange_nccl(lbm_handle_t h, const Plan &p, int64_t kl, int64_t ku, char *xbuf, int64_t eb) {
    // sends read / recvs write xbuf: order them after everything already queued on the main stream
    LBM_CUDA(h, cudaEventRecord(h->ev_ready, h->stream));
    LBM_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
    char *own = xbuf + p.left * eb;
    LBM_NCCL(h, g_nccl.group_start());
    if (h->rank > 0) {
        const int64_t cnt = ku < p.m_loc ? ku : p.m_loc;  // previous rank's right ghost = my first ku entries
        if (cnt > 0) LBM_NCCL(h, g_nccl.send(own, (size_t)(cnt * eb), 0 /*ncclInt8*/, h->rank - 1, h->nccl_comm, h->comm_stream));
        if (p.left > 0) LBM_NCCL(h, g_nccl.recv(xbuf, (size_t)(p.left * eb), 0, h->rank - 1, h->nccl_comm, h->comm_stream));
    }
    if (h->rank < h->world - 1) {
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

template <typename T> int gbmv_call(lbm_handle_t, int64_t, int64_t, int64_t, int64_t, T, const T *, int64_t, const T *, T, T *);
template <> int gbmv_call<double>(lbm_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, double alpha, const double *A,
                                  int64_t lda, const double *x, double beta, double *y) {
    return lbm_dgbmv(h, 'N', m, n, kl, ku, alpha, A, lda, x, beta, y);
}
template <> int gbmv_call<float>(lbm_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, float alpha, const float *A,
                                 int64_t lda, const float *x, float beta, float *y) {
    return lbm_sgbmv(h, 'N', m, n, kl, ku, alpha, A, lda, x, beta, y);
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
    if (p.m_loc == 0) return 0;
    if (h->world == 1) return apply_rows<T>(h, p, 0, p.m_loc, alpha, A, lda, xbuf, beta, y);
    if (!h->nccl_comm) return lbm_fail_arg(h, 1, "lbm_comm_init has not been called on this handle");
    if (p.m_loc < (kl > ku ? kl : ku)) return lbm_fail_arg(h, 2, "shard narrower than the halo max(kl,ku): use fewer ranks");
    int rc = exchange(h, p, kl, ku, reinterpret_cast<char *>(xbuf), (int)sizeof(T));
    if (rc) return rc;
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

}  // namespace

// internal (kron.cu): plan over `n` units of `unit_bytes` each + ghost exchange; the main stream does
// NOT wait -- the caller overlaps interior work and then waits on h->ev_halo.
int lbm_internal_exchange(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, void *buf, int64_t unit_bytes, int64_t *plan8) {
    const Plan p = make_plan(n, kl, ku, h->rank, h->world);
    plan8[0] = p.r0; plan8[1] = p.r1; plan8[2] = p.c0; plan8[3] = p.c1;
    plan8[4] = p.left; plan8[5] = p.right; plan8[6] = p.kl_loc; plan8[7] = p.ku_loc;
    if (h->world == 1) return 0;
    if (!h->nccl_comm) return lbm_fail_arg(h, 1, "lbm_comm_init has not been called on this handle");
    if (p.m_loc < (kl > ku ? kl : ku)) return lbm_fail_arg(h, 2, "shard narrower than the halo max(kl,ku): use fewer ranks");
    return exchange(h, p, kl, ku, reinterpret_cast<char *>(buf), unit_bytes);
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
    LBM_CUDA(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    LBM_CUDA(h, cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    LBM_CUDA(h, cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
    return 0;
}

// ---- peer-memory transport setup (all three are collective in spirit: every rank of the row partition calls them) ----
int lbm_peer_alloc(lbm_handle_t h, size_t slot_bytes, void *ipc_handle64) {
    if (!h) return -1;
    if (!ipc_handle64) return lbm_fail_arg(h, 3, "ipc_handle64 == NULL");
    if (slot_bytes == 0 || slot_bytes % 16) return lbm_fail_arg(h, 2, "slot_bytes must be a positive multiple of 16");
    lbm_device_guard g(h->device);
    lbm_peer_disconnect(h);
    const size_t total = PEER_FLAG_BYTES + 4 * slot_bytes;
    LBM_CUDA(h, cudaMalloc(&h->peer_local, total));
    LBM_CUDA(h, cudaMemset(h->peer_local, 0, total));
    LBM_CUDA(h, cudaDeviceSynchronize());
    h->peer_slot_bytes = slot_bytes;
    h->peer_epoch = 0;
    cudaIpcMemHandle_t ih;
    LBM_CUDA(h, cudaIpcGetMemHandle(&ih, h->peer_local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    memcpy(ipc_handle64, &ih, 64);
    return 0;
}

int lbm_peer_connect(lbm_handle_t h, const void *left_handle64, const void *right_handle64) {
    if (!h) return -1;
    if (!h->peer_local) return lbm_fail_arg(h, 1, "lbm_peer_alloc has not been called on this handle");
    if (!h->nccl_comm) return lbm_fail_arg(h, 1, "lbm_comm_init has not been called on this handle");
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
    return 0;
}

int lbm_peer_disconnect(lbm_handle_t h) {
    if (!h) return -1;
    lbm_device_guard g(h->device);
    if (h->peer_on || h->peer_local) cudaDeviceSynchronize();
    for (int s = 0; s < 2; ++s) {
        if (h->peer_remote[s]) cudaIpcCloseMemHandle(h->peer_remote[s]);
        h->peer_remote[s] = nullptr;
    }
    if (h->peer_local) cudaFree(h->peer_local);
    h->peer_local = nullptr;
    h->peer_on = 0;
    h->peer_slot_bytes = 0;
    return 0;
}

int lbm_peer_active(lbm_handle_t h) { return h && h->peer_on && h->tune.exchange_mode != 1 ? 1 : 0; }

int lbm_comm_destroy(lbm_handle_t h) {
    if (!h) return -1;
    lbm_peer_disconnect(h);
    if (h->nccl_comm) {
        lbm_device_guard g(h->device);
        cudaStreamSynchronize(h->comm_stream);
        g_nccl.destroy(h->nccl_comm);
        cudaEventDestroy(h->ev_ready);
        cudaEventDestroy(h->ev_halo);
        cudaStreamDestroy(h->comm_stream);
        h->nccl_comm = nullptr;
    }
    h->rank = 0;
    h->world = 1;
    return 0;
}

int lbm_halo_exchange(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, void *xbuf, int elem_bytes) {
    if (!h) return -1;
    if (h->world == 1) return 0;
    if (!h->nccl_comm) return lbm_fail_arg(h, 1, "lbm_comm_init has not been called on this handle");
    if (elem_bytes != 4 && elem_bytes != 8) return lbm_fail_arg(h, 6, "elem_bytes must be 4 or 8");
    lbm_device_guard g(h->device);
    const Plan p = make_plan(n, kl, ku, h->rank, h->world);
    int rc = exchange(h, p, kl, ku, reinterpret_cast<char *>(xbuf), elem_bytes);
    if (rc) return rc;
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
}
