// multi.cu -- SINGLE-PROCESS multi-device face of the row-sharded path ("multi-GPU is single-process
// multi-device, driven from one host thread"): what a Julia host needs to shard a BandedMatrix over the GPUs of one box
// without MPI.jl / Distributed.jl or any rendezvous of its own.
//
//   lbm_create_multi(ndev, devs)      one handle + stream per device, peer access enabled between row-neighbours,
//                                     mailbox blocks allocated and cross-linked by plain peer pointers (no IPC, no NCCL)
//   lbm_shard_create / _distribute    the row partition of an n x n banded operator: device r holds the column slice
//                                     [c0,c1) of the diagonal-major array (lbm_shard_plan), uploaded from ONE host array
//   lbm_svec_create / _distribute / _gather   vectors in the shard's ghost-cell layout
//   lbm_?gbmv_sharded_all             y <- alpha*A*x + beta*y on every device: one launch per device (ghost exchange fused
//                                     into the kernel, halo.cuh), all enqueued from the calling thread, nothing blocks
//   lbm_?gbmv_chain_sharded_all       mul!(y, ApplyMatrix(*, F_0..F_{nf-1}), x) with the intermediates' halos exchanged
//
// Every device-side mechanism is the one dist.cu uses for the process-per-GPU mode: the per-device handles carry
// rank / world / mailbox pointers, so lbm_?gbmv_sharded, lbm_?tridiag_mv_sharded and lbm_?kronsum2d_mv_sharded also work
// on the handles lbm_multi_handle() returns (with buffers from lbm_malloc on that handle).
#include "halo.cuh"

#include <stdlib.h>

struct lbm_multi_s {
    int ndev;
    lbm_handle_t *h;
    char err[512];
};

struct lbm_shard_s {
    lbm_multi_t m;
    int eb;                       // element bytes: 8 (Float64) or 4 (Float32)
    int64_t n, kl, ku, lda;
    void **slab;                  // per device: (c1-c0) columns x lda
    int64_t (*plan)[8];           // per device: lbm_shard_plan output
};

struct lbm_svec_s {
    lbm_shard_t s;
    void **buf;                   // per device: ghost-cell vector of c1-c0 elements (+ slack)
};

namespace {

int multi_fail(lbm_multi_t m, int code, const char *what) {
    if (m) snprintf(m->err, 512, "%s", what);
    return code;
}
int multi_fail_from(lbm_multi_t m, int r, int rc) {
    if (m) snprintf(m->err, 512, "device %d: %s", m->h[r]->device, lbm_last_error_string(m->h[r]));
    return rc;
}

inline int64_t mloc(const int64_t *p) { return p[1] - p[0]; }
inline int64_t nloc(const int64_t *p) { return p[1] > p[0] ? p[3] - p[2] : 0; }

}  // namespace

extern "C" {

int lbm_create_multi(lbm_multi_t *out, int ndev, const int *devices) {
    if (!out) return -1;
    *out = nullptr;
    if (ndev < 1) return -2;
    lbm_multi_t m = (lbm_multi_t)calloc(1, sizeof(lbm_multi_s));
    if (!m) return (int)cudaErrorMemoryAllocation;
    m->ndev = ndev;
    m->h = (lbm_handle_t *)calloc((size_t)ndev, sizeof(lbm_handle_t));
    int rc = 0;
    for (int r = 0; r < ndev && !rc; ++r) {
        rc = lbm_create(&m->h[r], devices ? devices[r] : r);
        if (rc) break;
        for (int q = 0; q < r; ++q)
            if (m->h[q]->device == m->h[r]->device) rc = -3;   // one shard per device
    }
    if (rc) {
        lbm_destroy_multi(m);
        return rc;
    }
    const size_t slot_bytes = (size_t)1 << 20;   // 131072 Float64 ghosts per side (sixteen 8192-point grid columns)
    for (int r = 0; r < ndev; ++r) {
        lbm_handle_t h = m->h[r];
        h->rank = r;
        h->world = ndev;
        h->has_comm = ndev > 1;
        if (ndev == 1) continue;
        lbm_device_guard g(h->device);
        if ((rc = lbm_internal_comm_resources(h))) break;
        for (int s = -1; s <= 1; s += 2) {
            const int q = r + s;
            if (q < 0 || q >= ndev) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, h->device, m->h[q]->device);
            if (!can) {
                snprintf(h->err, 512, "no peer access between devices %d and %d: the single-process mode needs NVLink/PCIe P2P",
                         h->device, m->h[q]->device);
                rc = (int)cudaErrorPeerAccessUnsupported;
                break;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(m->h[q]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                rc = lbm_fail_cuda(h, e, "cudaDeviceEnablePeerAccess");
                break;
            }
            cudaGetLastError();
        }
        if (rc) break;
        if ((rc = lbm_internal_peer_alloc_local(h, slot_bytes))) break;
    }
    if (!rc && ndev > 1) {
        for (int r = 0; r < ndev; ++r) {   // cross-link the mailboxes: plain peer pointers (UVA), no IPC handles
            lbm_handle_t h = m->h[r];
            h->peer_remote[0] = r > 0 ? m->h[r - 1]->peer_local : nullptr;
            h->peer_remote[1] = r < ndev - 1 ? m->h[r + 1]->peer_local : nullptr;
            h->peer_direct = 1;
            h->peer_on = 1;
        }
    }
    if (rc) {
        lbm_destroy_multi(m);
        return rc;
    }
    *out = m;
    return 0;
}

int lbm_destroy_multi(lbm_multi_t m) {
    if (!m) return 0;
    for (int r = 0; r < m->ndev; ++r)
        if (m->h[r]) {
            lbm_device_guard g(m->h[r]->device);
            cudaDeviceSynchronize();
        }
    for (int r = 0; r < m->ndev; ++r)
        if (m->h[r]) lbm_destroy(m->h[r]);
    free(m->h);
    free(m);
    return 0;
}

int lbm_multi_ndev(lbm_multi_t m) { return m ? m->ndev : -1; }
lbm_handle_t lbm_multi_handle(lbm_multi_t m, int r) { return (m && r >= 0 && r < m->ndev) ? m->h[r] : nullptr; }
const char *lbm_multi_last_error_string(lbm_multi_t m) { return m ? m->err : "NULL lbm_multi_t"; }

int lbm_multi_sync(lbm_multi_t m) {
    if (!m) return -1;
    int first = 0;
    for (int r = 0; r < m->ndev; ++r) {
        int rc = lbm_sync(m->h[r]);
        if (rc && !first) first = multi_fail_from(m, r, rc);
    }
    return first;
}

// ---- the sharded operator ---------------------------------------------------------------------------------------------
int lbm_shard_create(lbm_multi_t m, int elem_bytes, int64_t n, int64_t kl, int64_t ku, lbm_shard_t *out) {
    if (!m) return -1;
    if (!out) return multi_fail(m, -6, "out == NULL");
    *out = nullptr;
    if (elem_bytes != 4 && elem_bytes != 8) return multi_fail(m, -2, "elem_bytes must be 4 or 8");
    if (n < 0) return multi_fail(m, -3, "n < 0");
    if (kl < 0) return multi_fail(m, -4, "kl < 0");
    if (ku < 0) return multi_fail(m, -5, "ku < 0");
    lbm_shard_t s = (lbm_shard_t)calloc(1, sizeof(lbm_shard_s));
    s->m = m; s->eb = elem_bytes; s->n = n; s->kl = kl; s->ku = ku; s->lda = kl + ku + 1;
    s->slab = (void **)calloc((size_t)m->ndev, sizeof(void *));
    s->plan = (int64_t(*)[8])calloc((size_t)m->ndev, sizeof(int64_t[8]));
    int rc = 0;
    for (int r = 0; r < m->ndev && !rc; ++r) {
        rc = lbm_shard_plan(n, kl, ku, r, m->ndev, s->plan[r]);
        if (rc) {
            multi_fail(m, -4, "shards narrower than the halo max(kl,ku): use fewer devices");
            rc = -4;
            break;
        }
        rc = lbm_malloc(m->h[r], &s->slab[r], (size_t)(nloc(s->plan[r]) * s->lda + 4) * elem_bytes);
        if (rc) multi_fail_from(m, r, rc);
    }
    if (rc) {
        lbm_shard_destroy(s);
        return rc;
    }
    *out = s;
    return 0;
}

int lbm_shard_destroy(lbm_shard_t s) {
    if (!s) return 0;
    for (int r = 0; r < s->m->ndev; ++r)
        if (s->slab && s->slab[r]) lbm_free(s->m->h[r], s->slab[r]);
    free(s->slab);
    free(s->plan);
    free(s);
    return 0;
}

// A_host: the (kl+ku+1) x n diagonal-major array exactly as BandedMatrices stores it (column-major, leading dimension
// lda_host >= kl+ku+1; pinned or pageable).  Device r receives columns [c0,c1): ONE contiguous byte range when
// lda_host == kl+ku+1.  Copies are enqueued on every device's stream and overlap each other; lbm_multi_sync waits.
int lbm_shard_distribute(lbm_shard_t s, const void *A_host, int64_t lda_host) {
    if (!s) return -1;
    lbm_multi_t m = s->m;
    if (!A_host && s->n > 0) return multi_fail(m, -2, "A_host == NULL");
    if (lda_host < s->lda) return multi_fail(m, -3, "lda_host < kl+ku+1");
    for (int r = 0; r < m->ndev; ++r) {
        const int64_t *p = s->plan[r];
        const int64_t cols = nloc(p);
        if (cols == 0) continue;
        lbm_handle_t h = m->h[r];
        lbm_device_guard g(h->device);
        const char *src = (const char *)A_host + (size_t)(p[2] * lda_host) * s->eb;
        cudaError_t e;
        if (lda_host == s->lda)
            e = cudaMemcpyAsync(s->slab[r], src, (size_t)(cols * s->lda) * s->eb, cudaMemcpyHostToDevice, h->stream);
        else
            e = cudaMemcpy2DAsync(s->slab[r], (size_t)s->lda * s->eb, src, (size_t)lda_host * s->eb, (size_t)s->lda * s->eb,
                                  (size_t)cols, cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) return multi_fail_from(m, r, lbm_fail_cuda(h, e, "cudaMemcpyAsync(slab)"));
    }
    return 0;
}

// the device's slice of brand(n,n,kl,ku) generated in place from the counter-based stream (bit-identical to slicing the
// global matrix; no transfer) -- synthetic data at sizes that do not fit the host
int lbm_shard_fill_uniform(lbm_shard_t s, uint64_t seed) {
    if (!s) return -1;
    lbm_multi_t m = s->m;
    for (int r = 0; r < m->ndev; ++r) {
        const int64_t *p = s->plan[r];
        const int64_t cnt = nloc(p) * s->lda;
        if (cnt == 0) continue;
        int rc = s->eb == 8 ? lbm_dfill_uniform(m->h[r], seed, p[2] * s->lda, cnt, (double *)s->slab[r])
                            : lbm_sfill_uniform(m->h[r], seed, p[2] * s->lda, cnt, (float *)s->slab[r]);
        if (rc) return multi_fail_from(m, r, rc);
    }
    return 0;
}

// device pointer of rank r's slab and its plan {r0, r1, c0, c1, left, right, kl_loc, ku_loc}
int lbm_shard_slab(lbm_shard_t s, int r, void **dptr, int64_t *plan8) {
    if (!s) return -1;
    if (r < 0 || r >= s->m->ndev) return multi_fail(s->m, -2, "device index out of range");
    if (dptr) *dptr = s->slab[r];
    if (plan8) memcpy(plan8, s->plan[r], sizeof(int64_t[8]));
    return 0;
}

// ---- vectors in a shard's ghost-cell layout ---------------------------------------------------------------------------
int lbm_svec_create(lbm_shard_t s, lbm_svec_t *out) {
    if (!s) return -1;
    if (!out) return multi_fail(s->m, -2, "out == NULL");
    *out = nullptr;
    lbm_multi_t m = s->m;
    lbm_svec_t v = (lbm_svec_t)calloc(1, sizeof(lbm_svec_s));
    v->s = s;
    v->buf = (void **)calloc((size_t)m->ndev, sizeof(void *));
    for (int r = 0; r < m->ndev; ++r) {
        const size_t bytes = (size_t)(nloc(s->plan[r]) + 4) * s->eb;
        int rc = lbm_malloc(m->h[r], &v->buf[r], bytes);
        if (!rc) rc = lbm_memset0(m->h[r], v->buf[r], bytes);
        if (rc) {
            multi_fail_from(m, r, rc);
            lbm_svec_destroy(v);
            return rc;
        }
    }
    *out = v;
    return 0;
}

int lbm_svec_destroy(lbm_svec_t v) {
    if (!v) return 0;
    for (int r = 0; r < v->s->m->ndev; ++r)
        if (v->buf && v->buf[r]) lbm_free(v->s->m->h[r], v->buf[r]);
    free(v->buf);
    free(v);
    return 0;
}

// x_host (n elements) -> the owned block of every device's ghost-cell vector (ghost cells are filled by the exchange)
int lbm_svec_distribute(lbm_svec_t v, const void *x_host) {
    if (!v) return -1;
    lbm_shard_t s = v->s;
    if (!x_host && s->n > 0) return multi_fail(s->m, -2, "x_host == NULL");
    for (int r = 0; r < s->m->ndev; ++r) {
        const int64_t *p = s->plan[r];
        if (mloc(p) == 0) continue;
        int rc = lbm_memcpy_h2d(s->m->h[r], (char *)v->buf[r] + p[4] * s->eb, (const char *)x_host + p[0] * s->eb,
                                (size_t)mloc(p) * s->eb);
        if (rc) return multi_fail_from(s->m, r, rc);
    }
    return 0;
}

// the owned blocks -> y_host (n elements); enqueued on every device's stream, lbm_multi_sync waits
int lbm_svec_gather(lbm_svec_t v, void *y_host) {
    if (!v) return -1;
    lbm_shard_t s = v->s;
    if (!y_host && s->n > 0) return multi_fail(s->m, -2, "y_host == NULL");
    for (int r = 0; r < s->m->ndev; ++r) {
        const int64_t *p = s->plan[r];
        if (mloc(p) == 0) continue;
        int rc = lbm_memcpy_d2h(s->m->h[r], (char *)y_host + p[0] * s->eb, (const char *)v->buf[r] + p[4] * s->eb,
                                (size_t)mloc(p) * s->eb);
        if (rc) return multi_fail_from(s->m, r, rc);
    }
    return 0;
}

int lbm_svec_fill_uniform(lbm_svec_t v, uint64_t seed) {
    if (!v) return -1;
    lbm_shard_t s = v->s;
    for (int r = 0; r < s->m->ndev; ++r) {
        const int64_t *p = s->plan[r];
        if (mloc(p) == 0) continue;
        void *own = (char *)v->buf[r] + p[4] * s->eb;
        int rc = s->eb == 8 ? lbm_dfill_uniform(s->m->h[r], seed, p[0], mloc(p), (double *)own)
                            : lbm_sfill_uniform(s->m->h[r], seed, p[0], mloc(p), (float *)own);
        if (rc) return multi_fail_from(s->m, r, rc);
    }
    return 0;
}

// device pointer of rank r's ghost-cell buffer and of its owned block
int lbm_svec_buffer(lbm_svec_t v, int r, void **ghost_buf, void **owned) {
    if (!v) return -1;
    lbm_shard_t s = v->s;
    if (r < 0 || r >= s->m->ndev) return multi_fail(s->m, -2, "device index out of range");
    if (ghost_buf) *ghost_buf = v->buf[r];
    if (owned) *owned = (char *)v->buf[r] + s->plan[r][4] * s->eb;
    return 0;
}

}  // extern "C"

namespace {

template <typename T> struct Calls;
template <> struct Calls<double> {
    static int one(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, double a, const double *A, int64_t lda, double *x, double b, double *y) {
        return lbm_dgbmv_sharded(h, n, kl, ku, a, A, lda, x, b, y, 1);
    }
    static int chain(lbm_handle_t h, int nf, int64_t n, const int64_t *kl, const int64_t *ku, const double *const *A, const int64_t *lda,
                     double a, double *x, double b, double *y) {
        return lbm_dgbmv_chain_sharded(h, nf, n, kl, ku, A, lda, a, x, b, y);
    }
};
template <> struct Calls<float> {
    static int one(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, float a, const float *A, int64_t lda, float *x, float b, float *y) {
        return lbm_sgbmv_sharded(h, n, kl, ku, a, A, lda, x, b, y, 1);
    }
    static int chain(lbm_handle_t h, int nf, int64_t n, const int64_t *kl, const int64_t *ku, const float *const *A, const int64_t *lda,
                     float a, float *x, float b, float *y) {
        return lbm_sgbmv_chain_sharded(h, nf, n, kl, ku, A, lda, a, x, b, y);
    }
};

// x must live in the ghost-cell layout of the factor that reads it; y may live in any layout of the same n
template <typename T>
int check_operands(lbm_multi_t m, lbm_shard_t first_applied, lbm_svec_t x, lbm_svec_t y) {
    if (!x) return multi_fail(m, -3, "x == NULL");
    if (!y) return multi_fail(m, -5, "y == NULL");
    if (x->s->m != m || y->s->m != m) return multi_fail(m, -3, "operands belong to different lbm_multi_t objects");
    if ((int)sizeof(T) != first_applied->eb || x->s->eb != first_applied->eb || y->s->eb != first_applied->eb)
        return multi_fail(m, -1, "element type mismatch");
    if (x->s->n != first_applied->n || x->s->kl != first_applied->kl || x->s->ku != first_applied->ku)
        return multi_fail(m, -3, "DimensionMismatch: x is not in the ghost-cell layout (n, kl, ku) of the operator that reads it");
    if (y->s->n != first_applied->n) return multi_fail(m, -5, "DimensionMismatch: y has a different length");
    if (x == y) return multi_fail(m, -5, "x and y must be different vectors");
    return 0;
}

template <typename T>
int gbmv_all(lbm_shard_t A, T alpha, lbm_svec_t x, T beta, lbm_svec_t y) {
    if (!A) return -1;
    lbm_multi_t m = A->m;
    int rc = check_operands<T>(m, A, x, y);
    if (rc) return rc;
    for (int r = 0; r < m->ndev; ++r) {   // enqueue only: nothing here blocks, so every device's kernel gets launched
        T *yown = (T *)((char *)y->buf[r] + y->s->plan[r][4] * y->s->eb);
        rc = Calls<T>::one(m->h[r], A->n, A->kl, A->ku, alpha, (const T *)A->slab[r], A->lda, (T *)x->buf[r], beta, yown);
        if (rc) return multi_fail_from(m, r, rc);
    }
    return 0;
}

template <typename T>
int chain_all(int nf, const lbm_shard_t *F, T alpha, lbm_svec_t x, T beta, lbm_svec_t y) {
    if (nf < 1 || !F || !F[0]) return -1;
    lbm_multi_t m = F[0]->m;
    if (nf > 16) return multi_fail(m, -1, "at most 16 factors");
    for (int f = 0; f < nf; ++f) {
        if (!F[f] || F[f]->m != m) return multi_fail(m, -2, "factors belong to different lbm_multi_t objects");
        if (F[f]->n != F[0]->n) return multi_fail(m, -2, "DimensionMismatch: factors of different sizes");
        if (F[f]->eb != (int)sizeof(T)) return multi_fail(m, -2, "element type mismatch");
    }
    int rc = check_operands<T>(m, F[nf - 1], x, y);
    if (rc) return rc;
    // STAGE-major enqueue order (all devices of stage f before any device of stage f-1): a kernel waiting for its
    // neighbour's epoch then only ever waits for launches that are issued before the host touches anything that could
    // block on that device (a first-use module load, an allocation) -- rank-major order could deadlock the one host
    // thread against a spinning kernel under lazy module loading.
    const int64_t n = F[0]->n;
    T *v[64];
    T *tmp[64][2];
    if (m->ndev > 64) return multi_fail(m, -1, "at most 64 devices");
    for (int r = 0; r < m->ndev; ++r) {
        int64_t pitch = 4;
        for (int f = 0; f < nf; ++f) pitch = nloc(F[f]->plan[r]) > pitch ? nloc(F[f]->plan[r]) : pitch;
        pitch = (pitch + 31) / 32 * 32;
        void *scr = nullptr;
        if (nf > 1 && (rc = lbm_scratch(m->h[r], (size_t)(2 * pitch) * sizeof(T), &scr))) return multi_fail_from(m, r, rc);
        tmp[r][0] = (T *)scr;
        tmp[r][1] = (T *)scr + pitch;
        v[r] = (T *)x->buf[r];
    }
    for (int f = nf - 1; f >= 1; --f) {
        for (int r = 0; r < m->ndev; ++r) {
            T *t = tmp[r][f & 1];
            T *out = t + F[f - 1]->plan[r][4];   // owned block inside the CONSUMER's ghost-cell layout
            rc = Calls<T>::one(m->h[r], n, F[f]->kl, F[f]->ku, (T)1, (const T *)F[f]->slab[r], F[f]->lda, v[r], (T)0, out);
            if (rc) return multi_fail_from(m, r, rc);
            v[r] = t;
        }
    }
    for (int r = 0; r < m->ndev; ++r) {
        T *yown = (T *)((char *)y->buf[r] + y->s->plan[r][4] * y->s->eb);
        rc = Calls<T>::one(m->h[r], n, F[0]->kl, F[0]->ku, alpha, (const T *)F[0]->slab[r], F[0]->lda, v[r], beta, yown);
        if (rc) return multi_fail_from(m, r, rc);
    }
    return 0;
}

}  // namespace

extern "C" {
int lbm_dgbmv_sharded_all(lbm_shard_t A, double alpha, lbm_svec_t x, double beta, lbm_svec_t y) { return gbmv_all<double>(A, alpha, x, beta, y); }
int lbm_sgbmv_sharded_all(lbm_shard_t A, float alpha, lbm_svec_t x, float beta, lbm_svec_t y) { return gbmv_all<float>(A, alpha, x, beta, y); }
int lbm_dgbmv_chain_sharded_all(int nf, const lbm_shard_t *F, double alpha, lbm_svec_t x, double beta, lbm_svec_t y) {
    return chain_all<double>(nf, F, alpha, x, beta, y);
}
int lbm_sgbmv_chain_sharded_all(int nf, const lbm_shard_t *F, float alpha, lbm_svec_t x, float beta, lbm_svec_t y) {
    return chain_all<float>(nf, F, alpha, x, beta, y);
}
}
