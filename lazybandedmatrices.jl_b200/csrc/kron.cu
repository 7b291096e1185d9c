// kron.cu -- block-Kronecker operators applied WITHOUT forming them.
//
//   kronsum2d:  y = alpha*(kron(A,I) + kron(I,B)) x + beta*y = alpha*vec(X A^T + B X) + beta*y
//   kron2d:     y = alpha*kron(A,B) x + beta*y             = alpha*vec(B X A^T)     + beta*y
// with x = vec(X), X n2 x n1 column-major.  This is BlockKron(A,I)+BlockKron(I,B) -- the 2-D
// Laplacian of /root/reference/test/test_blockkron.jl:296-306 -- via the identity the reference
// itself applies at src/blockkron.jl:272-278 and :440.  The reference's own route materialises
// a 9-values-per-column BandedBlockBandedMatrix (4.8 GB at 8192^2); here the only HBM traffic
// is X in, Y out, plus the two tiny band arrays.
//
//  * kron_ring_kernel (narrow bands, aligned grid -- the 2-D Laplacian case): a CTA owns one strip
//    of TR rows and walks column chunks of G columns; each chunk's halo'd X tile arrives as one
//    1-D bulk TMA copy per grid column into a multi-stage smem ring (mbarrier-guarded).  B's
//    coefficients for the thread's rows stay in registers for the whole sweep, A's per-column
//    coefficients are warp-uniform loads, Y is written coalesced.  Edge tiles predicate by INDEX.
//  * kron_tile_kernel (any bands / any alignment): a CTA owns a TR x TC tile of Y, stages the
//    (TR+klb+kub) x (TC+kla+kua) halo'd tile of X in shared memory (zero outside the grid) and
//    the tile's coefficient rows of A and B with out-of-matrix slots masked to zero BY INDEX.
#include "halo.cuh"


namespace {

template <typename T>
struct KronArgs {
    int64_t n1, n2;
    int kla, kua, klb, kub;
    int64_t lda, ldb;
    const T *A;
    const T *B;
    T alpha, beta;
    const T *x;
    T *y;
    int tr, tc;       // tile rows / cols
    int pitch;        // smem pitch of the X tile (rows incl. halo, padded)
    int offCA, offCB; // smem offsets (elements)
    int useA, useB;   // which terms to apply
    int64_t batch_stride;  // elements between consecutive (n2 x n1) slices (blockIdx.z)
    // sharded use: X has n1_in columns, output column c reads input column c + xoff; only output
    // columns [c_begin, c_end) are produced.  Unsharded: n1_in = n1, xoff = 0, range = [0, n1).
    int64_t n1_in, xoff, c_begin, c_end;
};

template <typename T>
__global__ void __launch_bounds__(256) kron_tile_kernel(const KronArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sX = reinterpret_cast<T *>(smem_raw);
    T *sCA = sX + a.offCA;  // [tc][nba]  A[c, c-kla+t]
    T *sCB = sX + a.offCB;  // [tr][nbb]  B[r, r-klb+t]
    const int tid = threadIdx.x;
    const int nba = a.kla + a.kua + 1, nbb = a.klb + a.kub + 1;
    const int64_t r0 = (int64_t)blockIdx.x * a.tr;
    const int64_t c0 = a.c_begin + (int64_t)blockIdx.y * a.tc;
    const T *xs = a.x + (int64_t)blockIdx.z * a.batch_stride;
    T *ys = a.y + (int64_t)blockIdx.z * a.batch_stride;
    const int trn = (int)(r0 + a.tr < a.n2 ? a.tr : a.n2 - r0);
    const int tcn = (int)(c0 + a.tc < a.c_end ? a.tc : a.c_end - c0);
    const int hl = a.useB ? a.klb : 0, hu = a.useB ? a.kub : 0;   // row halo
    const int wl = a.useA ? a.kla : 0, wu = a.useA ? a.kua : 0;   // column halo
    const int rows_h = trn + hl + hu;
    const int cols_h = tcn + wl + wu;

    // X tile with halo; zero outside the grid
    for (int e = tid; e < rows_h * cols_h; e += 256) {
        const int cc = e / rows_h;
        const int rr = e - cc * rows_h;
        const int64_t r = r0 - hl + rr;
        const int64_t c = c0 + a.xoff - wl + cc;  // input column
        T v = 0;
        if (r >= 0 && r < a.n2 && c >= 0 && c < a.n1_in) v = xs[r + a.n2 * c];
        sX[rr + a.pitch * cc] = v;
    }
    if (a.useA)
        for (int e = tid; e < tcn * nba; e += 256) {
            const int cc = e / nba;
            const int t = e - cc * nba;
            const int64_t c = c0 + cc;
            const int64_t q = c + a.xoff - a.kla + t;  // input column; A[c,q] sits at band row kua+kla-t of slab column q
            sCA[e] = (q >= 0 && q < a.n1_in) ? a.A[(a.kua + a.kla - t) + a.lda * q] : (T)0;
        }
    if (a.useB)
        for (int e = tid; e < trn * nbb; e += 256) {
            const int rr = e / nbb;
            const int t = e - rr * nbb;
            const int64_t r = r0 + rr;
            const int64_t p = r - a.klb + t;  // B[r,p] = Bd[(kub + r - p) + ldb*p]
            sCB[e] = (p >= 0 && p < a.n2) ? a.B[(a.kub + r - p) + a.ldb * p] : (T)0;
        }
    __syncthreads();

    for (int rr = tid; rr < trn; rr += 256) {
        const int64_t r = r0 + rr;
        for (int cc = 0; cc < tcn; ++cc) {
            T acc = 0;
            if (a.useA) {
                const T *xa = sX + (rr + hl) + a.pitch * cc;  // X[r, c-kla+t]
                const T *ca = sCA + cc * nba;
                for (int t = 0; t < nba; ++t) acc = fma(ca[t], xa[a.pitch * t], acc);
            }
            if (a.useB) {
                const T *xb = sX + rr + a.pitch * (cc + wl);  // X[r-klb+t, c]
                const T *cb = sCB + rr * nbb;
                for (int t = 0; t < nbb; ++t) acc = fma(cb[t], xb[t], acc);
            }
            const int64_t o = r + a.n2 * (c0 + cc);
            ys[o] = (a.beta == (T)0) ? a.alpha * acc : fma(a.beta, ys[o], a.alpha * acc);
        }
    }
}


// ------------------------------------------------------------------------------------------
// TMA-ring kernel for narrow bands
// ------------------------------------------------------------------------------------------
using namespace lbm;

template <typename T>
struct KronRingArgs {
    int64_t n1, n2;
    int kla, kua, klb, kub;
    int64_t lda, ldb;
    const T *A;
    const T *B;
    T alpha, beta;
    const T *x;
    T *y;
    int tr;        // rows per strip (256 * RPT)
    int g;         // output columns per chunk
    int hp;        // row halo padded to a 16-byte multiple
    int seg;       // smem elements per staged grid column = tr + 2*hp
    int stages;
    int64_t nchunks;
    int64_t n1_in, xoff, c_begin, c_end;  // see KronArgs
    int prod;      // 0: X*A^T + B*X (Kronecker sum); 1: B*X*A^T (Kronecker product, one pass)
    int64_t batch_stride;   // blockIdx.z walks independent slices (x, y advance by this many elements): 3-D mode products
};

// Ghost columns of one stage: wait for the strip's flag, then overwrite rows [r0, r0+rows) of every ghost column the stage
// holds from the local mailbox (only those rows are ever read: the A-term X[r, c+-t]; the B-term reads owned columns).
template <typename T>
__device__ __noinline__ void kron_halo_patch(const HaloFuse &hf, T *st, int64_t ci0, int scols, int64_t n1_in, int seg, int rowoff,
                                             int64_t r0, int rows, int64_t n2, int strip) {
    const int tid = threadIdx.x;
    const int64_t sa = ci0 > 0 ? ci0 : 0;
    const int64_t sb = ci0 + scols < n1_in ? ci0 + scols : n1_in;
    const bool need0 = hf.ghost_end[0] > hf.ghost_begin[0] && sa < hf.ghost_end[0] && sb > hf.ghost_begin[0];
    const bool need1 = hf.ghost_end[1] > hf.ghost_begin[1] && sa < hf.ghost_end[1] && sb > hf.ghost_begin[1];
    if (!(need0 || need1)) return;   // CTA-uniform
    if (tid == 0 && need0) halo_wait_flag(hf.pull_flag[0] + strip, hf.epoch, hf.timeout_ns, hf.status);
    if (tid == 1 && need1) halo_wait_flag(hf.pull_flag[1] + strip, hf.epoch, hf.timeout_ns, hf.status);
    __syncthreads();
    if (need0) {
        const int64_t ga = sa > hf.ghost_begin[0] ? sa : hf.ghost_begin[0], gb = sb < hf.ghost_end[0] ? sb : hf.ghost_end[0];
        const T *mb = reinterpret_cast<const T *>(hf.pull_src[0]);
        for (int64_t c = ga; c < gb; ++c)
            for (int r = tid; r < rows; r += 256) st[(c - ci0) * seg + rowoff + r] = __ldcg(mb + (c - hf.ghost_begin[0]) * n2 + r0 + r);
    }
    if (need1) {
        const int64_t ga = sa > hf.ghost_begin[1] ? sa : hf.ghost_begin[1], gb = sb < hf.ghost_end[1] ? sb : hf.ghost_end[1];
        const T *mb = reinterpret_cast<const T *>(hf.pull_src[1]);
        for (int64_t c = ga; c < gb; ++c)
            for (int r = tid; r < rows; r += 256) st[(c - ci0) * seg + rowoff + r] = __ldcg(mb + (c - hf.ghost_begin[1]) * n2 + r0 + r);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
}

// HALO: grid-column-sharded apply in ONE launch (halo.cuh).  CTA (0,0) pushes this rank's boundary grid columns to the
// neighbours' mailboxes; column chunks are walked rotated by one so the first / last chunk (the only ones whose stage
// holds ghost columns) come last, and those chunks overwrite the ghost columns of their stage -- rows [rs,re) of this
// CTA's strip only -- in shared memory from the local mailbox.
template <typename T, int NBA, int NBB, int RPT, bool HALO>
__global__ void __launch_bounds__(256, HALO ? 2 : 0) kron_ring_kernel(const KronRingArgs<T> a, const __grid_constant__ HaloFuse hf) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    T *sbase = reinterpret_cast<T *>(smem_raw + 128);
    const int tid = threadIdx.x;
    const int S = a.stages;
    const int scols = a.g + a.kla + a.kua;
    const int64_t stage_elems = (int64_t)scols * a.seg;
    const int64_t r0 = (int64_t)blockIdx.x * a.tr;
    const T *const xz = a.x + (int64_t)blockIdx.z * a.batch_stride;
    T *const yz = a.y + (int64_t)blockIdx.z * a.batch_stride;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();

    const int64_t step = gridDim.y;
    const int64_t mine = (int64_t)blockIdx.y < a.nchunks ? (a.nchunks - blockIdx.y + step - 1) / step : 0;
    // logical chunk L = blockIdx.y + kk*step; physical chunk = L+1 (mod nchunks) when HALO.  All code below works on
    // physical chunk numbers: chunk_of(kk).
    // HALO: the rotation is chosen so that the two ghost-touching chunks (physical 0 and nchunks-1) are the LAST chunks of
    // the two CTAs with the highest blockIdx.y -- the ones that own one chunk fewer than the others whenever the chunks do
    // not divide evenly -- and the push / flag duty goes to a third such CTA: the exchange work then sits on CTAs that
    // would otherwise idle at the end instead of on the ones that finish last.
    int64_t rot = 1;      // physical = (logical + rot) mod nchunks
    int ypush = 0;
    if (HALO && !(hf.on & 2) && step >= 2 && a.nchunks >= step) {
        const int64_t m1 = (a.nchunks - (step - 1) + step - 1) / step, m2 = (a.nchunks - (step - 2) + step - 1) / step;
        if (m1 == m2) {
            const int64_t L1 = (step - 1) + (m1 - 1) * step;   // last logical chunk of CTA step-1; CTA step-2 ends at L1 - 1
            rot = a.nchunks - L1;
            if (step >= 3) ypush = (int)step - 3;
        }
    }
    auto chunk_of = [&](int64_t kk) -> int64_t {
        const int64_t L = blockIdx.y + kk * step;
        if (!(HALO && !(hf.on & 2))) return L;
        const int64_t q = L + rot;
        return q >= a.nchunks ? q - a.nchunks : q;
    };
    // trailing chunks of this CTA whose stage holds ghost columns (see the two loops at the end)
    int64_t ntail = 0;
    if (HALO && !(hf.on & 4)) {
        while (ntail < mine && ntail < 2) {
            const int64_t q = chunk_of(mine - 1 - ntail);
            if (q == 0 || q + 1 == a.nchunks) ++ntail;
            else break;
        }
    }
    // rows staged per grid column: [rs, re) (both 16-byte aligned: n2 % EPV == 0)
    const int64_t rs = r0 - a.hp > 0 ? r0 - a.hp : 0;
    const int64_t re = r0 + a.tr + a.hp < a.n2 ? r0 + a.tr + a.hp : a.n2;
    const uint32_t col_bytes = (uint32_t)((re - rs) * (int64_t)sizeof(T));
    const int roff = (int)(rs - (r0 - a.hp));  // smem row offset of rs

    // Warp 0 issues a chunk: lane 0 arms the byte count, the lanes issue one per-column bulk copy
    // each IN PARALLEL (one instruction for up to 32 copies), lane 0 arrives.
    auto issue = [&](int64_t q, int s) {  // all 32 lanes of warp 0
        const int lane = tid & 31;
        const int64_t c0 = a.c_begin + q * a.g;
        const int64_t ci0 = c0 + a.xoff - a.kla;  // first INPUT column of the stage
        int64_t ca = ci0 > 0 ? ci0 : 0;
        int64_t cb = ci0 + a.g + a.kla + a.kua < a.n1_in ? ci0 + a.g + a.kla + a.kua : a.n1_in;
        if (c0 + a.g > a.c_end) {  // last (short) chunk of the range: no columns beyond its halo
            const int64_t lim = a.c_end + a.xoff + a.kua;
            if (cb > lim) cb = lim;
        }
        if (cb < ca) cb = ca;
        T *st = sbase + (int64_t)s * stage_elems;
        if (lane == 0) mbar_add_tx(&bars[s], (uint32_t)(cb - ca) * col_bytes);
        __syncwarp();
        for (int64_t c = ca + lane; c < cb; c += 32)
            bulk_g2s(st + (c - ci0) * a.seg + roff, xz + c * a.n2 + rs, col_bytes, &bars[s]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[s]);
    };
    if (tid < 32) {
        const int64_t pro = mine < S ? mine : S;
        for (int64_t k = 0; k < pro; ++k) issue(chunk_of(k), (int)k);
    }
    // STRIP-WISE push (rides on the latency of the prologue copies): the CTA (strip, 0) sends rows [r0, r0+tr) of this
    // rank's boundary grid columns -- all the neighbour's CTAs of the same strip need -- and publishes the epoch in the
    // strip's own flag word, so no CTA ever moves more than a few KB and nobody waits for rows it does not use.
    if (HALO && (int)blockIdx.y == ypush) {
        const int64_t rows = (r0 + a.tr < a.n2 ? r0 + a.tr : a.n2) - r0;
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            if (!hf.push_flag[sd]) continue;
            const int64_t colbytes = a.n2 * (int64_t)sizeof(T);
            const int64_t ncols = hf.push_bytes[sd] / colbytes;
            for (int64_t c = 0; c < ncols; ++c)
                halo_copy(hf.push_src[sd] + c * colbytes + r0 * (int64_t)sizeof(T), hf.push_dst[sd] + c * colbytes + r0 * (int64_t)sizeof(T),
                          rows * (int64_t)sizeof(T), tid, 256);
        }
        // The flags are published AFTER this CTA's first chunk (below): by then the peer stores above have long been
        // acknowledged, so the system-scope fence costs nothing; fencing here held all 256 threads for a NVLink round trip
        // (~3 us of a 28 us kernel, on exactly the CTAs that then finish last).  The neighbour needs the ghosts only for
        // ITS last chunks.  Exception: a CTA whose FIRST chunk already waits for the neighbour (slabs of one or two chunks)
        // must publish before it waits, or the pair would wait for each other.
        if (mine <= ntail) {
            __threadfence_system();
            __syncthreads();
            if (tid == 0 && hf.push_flag[0]) st_release_sys(hf.push_flag[0] + blockIdx.x, hf.epoch);
            if (tid == 1 && hf.push_flag[1]) st_release_sys(hf.push_flag[1] + blockIdx.x, hf.epoch);
        }
    }

    // A coefficients of a chunk's columns, ca[cc][t] = A[c, c-kla+t] (zero outside the matrix):
    // one element per thread, prefetched into a register one chunk ahead and parked in a
    // double-buffered smem table, so the column loop never waits on a global load.
    T *sCA = sbase + (int64_t)S * stage_elems;  // [2][g*NBA]
    const int ncoef = a.g * NBA;
    auto load_coef = [&](int64_t q) -> T {
        if (tid >= ncoef) return (T)0;
        const int cc = tid / NBA, t = tid - cc * NBA;
        const int64_t c = a.c_begin + q * a.g + cc;
        const int64_t qq = c + a.xoff - a.kla + t;  // input column
        return (c < a.c_end && qq >= 0 && qq < a.n1_in) ? a.A[(a.kua + a.kla - t) + a.lda * qq] : (T)0;
    };
    if (mine > 0 && tid < ncoef) sCA[tid] = load_coef(chunk_of(0));

    // B coefficients of this thread's rows: cb[k][t] = B[r, r-klb+t], zero outside the matrix
    T cbv[RPT][NBB];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int64_t r = r0 + tid + 256 * k;
#pragma unroll
        for (int t = 0; t < NBB; ++t) {
            const int64_t p = r - a.klb + t;
            cbv[k][t] = (r < a.n2 && p >= 0 && p < a.n2) ? a.B[(a.kub + a.klb - t) + a.ldb * p] : (T)0;
        }
    }
    // Halo rows outside the grid (first / last strip) are never written by the bulk copies:
    // zero them once in every stage.  With B's out-of-matrix coefficients already zero this makes
    // the unmasked interior path exact for edge strips too (0 * 0, never 0 * garbage), so only a
    // PARTIAL last strip or the first/last column chunk needs the predicated path.
    {
        const int rows_in = (int)(re - rs);
        const int nz = a.seg - rows_in;  // rows [0, roff) and [roff + rows_in, seg)
        for (int e = tid; e < S * scols * nz; e += 256) {
            const int col = e / nz, z = e - col * nz;
            const int row = z < roff ? z : rows_in + z;
            sbase[(int64_t)col * a.seg + row] = (T)0;
        }
    }
    const bool row_edge = r0 + a.tr > a.n2;  // partial strip: some threads own no row
    __syncthreads();  // sCA[0] and the zeroed halo rows visible

    // With the exchange fused in, the chunks whose stage holds ghost columns are the LAST one or two of the CTAs that own
    // them (rotation above).  They run in a second copy of the chunk body, so the streaming loop proper carries no exchange
    // code at all: with the patch call inside the one loop the allocator spilled 64 bytes in it (19 % slower at every size).
    auto do_chunk = [&](const int64_t kk, const bool patch) {
        const int s = (int)(kk % S);
        const uint32_t parity = (uint32_t)((kk / S) & 1);
        const int64_t q = chunk_of(kk);
        const int64_t c0 = a.c_begin + q * a.g;
        const int gq = (int)(c0 + a.g < a.c_end ? a.g : a.c_end - c0);
        const T *st = sbase + (int64_t)s * stage_elems;
        const T *ca = sCA + (kk & 1) * ncoef;
        const bool edge = row_edge || (c0 + a.xoff - a.kla < 0) || (c0 + a.xoff + gq + a.kua > a.n1_in);
        T cnext = (T)0;
        if (kk + 1 < mine) cnext = load_coef(chunk_of(kk + 1));  // in flight during this chunk's compute
        mbar_wait(&bars[s], parity);
        // only the first / last chunk of the rank's column range can hold ghost columns: one cheap uniform test here, the
        // wait + patch itself out of line so that it does not touch the register allocation of the streaming loop
        if (HALO && patch)
            kron_halo_patch<T>(hf, const_cast<T *>(st), c0 + a.xoff - a.kla, a.g + a.kla + a.kua, a.n1_in, a.seg, roff + (int)(r0 - rs), r0,
                               (int)((r0 + a.tr < a.n2 ? r0 + a.tr : a.n2) - r0), a.n2, (int)blockIdx.x);
        const int seg = a.seg;
        const T *xrow = st + tid + a.hp;                 // X[r0+tid, c0-kla] of this stage
        T *yp = yz + (r0 + tid) + a.n2 * c0;              // Y[r0+tid, c0]; += n2 per column
        if (!edge) {
            // interior chunk: no index tests, 32-bit smem offsets, one pointer bump per column
#pragma unroll 4
            for (int cc = 0; cc < gq; ++cc) {
                T cav[NBA];
#pragma unroll
                for (int t = 0; t < NBA; ++t) cav[t] = ca[cc * NBA + t];
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const T *xa = xrow + cc * seg + 256 * k;                // X[r, c-kla+t] at xa[t*seg]
                    const T *xb = xa + a.kla * seg - a.klb;                 // X[r-klb+t, c] at xb[t]
                    T acc = 0;
                    if (a.prod) {
                        // (B*X)[r, c-kla+tA] for each of the NBA columns, combined with A's row: ((B X) A^T)[r, c]
#pragma unroll
                        for (int tA = 0; tA < NBA; ++tA) {
                            const T *xc = xa + tA * seg - a.klb;
                            T bx = 0;
#pragma unroll
                            for (int tB = 0; tB < NBB; ++tB) bx = fma(cbv[k][tB], xc[tB], bx);
                            acc = fma(cav[tA], bx, acc);
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < NBA; ++t) acc = fma(cav[t], xa[t * seg], acc);
#pragma unroll
                        for (int t = 0; t < NBB; ++t) acc = fma(cbv[k][t], xb[t], acc);
                    }
                    T *yo = yp + 256 * k;
                    *yo = (a.beta == (T)0) ? a.alpha * acc : fma(a.beta, *yo, a.alpha * acc);
                }
                yp += a.n2;
            }
        } else {
            for (int cc = 0; cc < gq; ++cc) {
                const int64_t c = c0 + cc;
                T cav[NBA];
#pragma unroll
                for (int t = 0; t < NBA; ++t) cav[t] = ca[cc * NBA + t];
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const int64_t r = r0 + tid + 256 * k;
                    if (r < a.n2) {
                        const T *xa = xrow + cc * seg + 256 * k;
                        const T *xb = xa + a.kla * seg - a.klb;
                        T acc = 0;
                        if (a.prod) {
#pragma unroll
                            for (int tA = 0; tA < NBA; ++tA) {
                                const int64_t qq = c + a.xoff - a.kla + tA;
                                if (qq >= 0 && qq < a.n1_in) {
                                    const T *xc = xa + tA * seg - a.klb;
                                    T bx = 0;
#pragma unroll
                                    for (int tB = 0; tB < NBB; ++tB) {
                                        const int64_t p = r - a.klb + tB;
                                        if (p >= 0 && p < a.n2) bx = fma(cbv[k][tB], xc[tB], bx);
                                    }
                                    acc = fma(cav[tA], bx, acc);
                                }
                            }
                        } else {
#pragma unroll
                        for (int t = 0; t < NBA; ++t) {
                            const int64_t qq = c + a.xoff - a.kla + t;
                            if (qq >= 0 && qq < a.n1_in) acc = fma(cav[t], xa[t * seg], acc);
                        }
#pragma unroll
                        for (int t = 0; t < NBB; ++t) {
                            const int64_t p = r - a.klb + t;
                            if (p >= 0 && p < a.n2) acc = fma(cbv[k][t], xb[t], acc);
                        }
                        }
                        T *yo = yp + 256 * k;
                        *yo = (a.beta == (T)0) ? a.alpha * acc : fma(a.beta, *yo, a.alpha * acc);
                    }
                }
                yp += a.n2;
            }
        }
        if (kk + 1 < mine && tid < ncoef) sCA[((kk + 1) & 1) * ncoef + tid] = cnext;
        __syncthreads();
        if (tid < 32 && kk + S < mine) issue(chunk_of(kk + S), s);
        if (HALO && kk == 0 && (int)blockIdx.y == ypush && mine > ntail) {
            // every thread's payload stores precede the barrier above; one thread per side fences at system scope and
            // releases the strip's flag word (warp 2: not the warp that issues the bulk copies)
            if (tid == 64 && hf.push_flag[0]) {
                __threadfence_system();
                st_release_sys(hf.push_flag[0] + blockIdx.x, hf.epoch);
            }
            if (tid == 65 && hf.push_flag[1]) {
                __threadfence_system();
                st_release_sys(hf.push_flag[1] + blockIdx.x, hf.epoch);
            }
        }
    };
    for (int64_t kk = 0; kk < mine - ntail; ++kk) do_chunk(kk, false);
    if (HALO)
        for (int64_t kk = mine - ntail; kk < mine; ++kk) do_chunk(kk, true);
    if (HALO && (int)blockIdx.y == ypush) halo_final_wait(hf, tid, blockIdx.x);
}

// Chunk width, ring depth and CTAs per SM are chosen HERE, per instantiation, because they depend on how many CTAs the
// register file admits (wider bands and the fused-exchange variant need > 85 registers: two resident CTAs, not three; a grid
// sized for three would run a second partial wave).
// Defaults from benchmarks/kron_ab.py (graph replay, alternating rounds, 8192-row slabs of 512 ... 8192 grid columns,
// profiles/kron_ab_r02.jsonl): 512-row strips; with three resident CTAs ~25 KB stages (4 columns + halo for the f64
// Laplacian) in a 2-stage ring are the best or within 1.5 % of it at every slab width from 1024 columns up (8192: 195 ->
// 174.5 us, 94 % of the copy ceiling; 1024, one rank of eight: 37.8 -> 27.1 us); with two, twice the stage.  4 CTAs/SM is
// 25-50 % slower at every width, as are 16-column chunks.  Slabs of <= 512 columns have too few chunks per CTA for a ring:
// one 8-column stage, 4 CTAs/SM.
template <typename T, int NBA, int NBB, int RPT, bool HALO>
int kron_ring_launch4(lbm_handle_t h, KronRingArgs<T> a, int64_t batch, const HaloFuse &hf) {
    static std::atomic<int> configured_dev_mask{0};
    static std::atomic<int> reg_ctas{0};
    auto kern = kron_ring_kernel<T, NBA, NBB, RPT, HALO>;
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
        cudaFuncAttributes fa;
        LBM_CUDA(h, cudaFuncGetAttributes(&fa, kern));
        const int per_cta = ((fa.numRegs + 7) / 8 * 8) * 256;
        reg_ctas = per_cta > 0 ? 65536 / per_cta : 1;
        configured_dev_mask |= (1 << h->device);
    }
    const int64_t occ = reg_ctas < 1 ? 1 : (int64_t)reg_ctas;
    const int64_t ncols = a.c_end - a.c_begin, nh = a.kla + a.kua;
    int64_t cps = h->tune.kron_ctas_per_sm > 0 ? h->tune.kron_ctas_per_sm : (ncols <= 512 ? 4 : 3);
    if (cps > occ) cps = occ;
    if (cps < 1) cps = 1;
    if (cps > 8) cps = 8;
    int64_t g = h->tune.kron_tc;
    if (g <= 0) {
        const int64_t target = cps >= 3 ? 24960 : cps == 2 ? 41600 : 100000;   // bytes per stage
        g = target / ((int64_t)a.seg * (int64_t)sizeof(T)) - nh;
        g = g >= 16 ? 16 : g >= 8 ? 8 : 4;
        if (ncols <= 512 && g < 16) g *= 2;
    }
    if (g > ncols) g = ncols;
    if (g * NBA > 256) g = 256 / NBA;  // one A coefficient per thread per chunk
    int64_t stages = h->tune.kron_stages > 0 ? h->tune.kron_stages : (ncols <= 512 ? 1 : 2);
    const int64_t sm_smem = 227 * 1024;
    auto stage_bytes = [&](int64_t gg) { return (gg + nh) * (int64_t)a.seg * (int64_t)sizeof(T); };
    while (g > 2 && stages * stage_bytes(g) + 4096 > sm_smem) g /= 2;
    while (stages > 1 && cps * (stages * stage_bytes(g) + 4096) > sm_smem) --stages;
    while (cps > 1 && cps * (stages * stage_bytes(g) + 4096) > sm_smem) --cps;
    a.g = (int)g;
    a.nchunks = (ncols + g - 1) / g;
    a.stages = (int)stages;
    const int64_t nstrips = (a.n2 + a.tr - 1) / a.tr;
    int64_t phases = ((int64_t)h->sm_count * cps) / (nstrips * batch);  // floor: every CTA resident, one wave
    if (phases > a.nchunks) phases = a.nchunks;
    if (phases > 65535) phases = 65535;
    if (phases < 1) phases = 1;   // batched slices: one CTA per (strip, slice) walks all its chunks, several waves
    dim3 grid((unsigned)nstrips, (unsigned)phases, (unsigned)batch);
    const size_t smem = (size_t)(128 + stages * stage_bytes(g) + 2 * g * NBA * (int64_t)sizeof(T));
    LBM_CUDA(h, lbm_launch_pdl(h, kern, grid, dim3(256), smem, a, hf));
    LBM_LAUNCH_CHECK(h, "kron_ring_kernel");
    return 0;
}

template <typename T, int NBA, int NBB, int RPT>
int kron_ring_launch3(lbm_handle_t h, const KronRingArgs<T> &a, int64_t batch, const HaloFuse *hf) {
    if (hf) return kron_ring_launch4<T, NBA, NBB, RPT, true>(h, a, batch, *hf);
    HaloFuse none;
    none.on = 0;
    return kron_ring_launch4<T, NBA, NBB, RPT, false>(h, a, batch, none);
}

template <typename T, int NBA, int NBB>
int kron_ring_launch2(lbm_handle_t h, const KronRingArgs<T> &a, int64_t batch, const HaloFuse *hf) {
    switch (a.tr / 256) {
        case 1: return kron_ring_launch3<T, NBA, NBB, 1>(h, a, batch, hf);
        case 2: return kron_ring_launch3<T, NBA, NBB, 2>(h, a, batch, hf);
        default: return kron_ring_launch3<T, NBA, NBB, 4>(h, a, batch, hf);
    }
}

// returns true (and the launch status in *rc) when the ring kernel took the call
// rows per strip of the ring kernel (tuning knob kron_tr: 256 / 512 / 1024); also the unit of the fused exchange's flag words
static inline int64_t kron_ring_tr(lbm_handle_t h) {
    const int64_t t = h->tune.kron_tr > 0 ? h->tune.kron_tr : 512;
    return t >= 1024 ? 1024 : t >= 512 ? 512 : 256;
}

template <typename T>
bool kron_ring_try(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const T *A, int64_t lda,
                   int64_t klb, int64_t kub, const T *B, int64_t ldb, T alpha, const T *x, T beta, T *y, int *rc,
                   int64_t n1_in = -1, int64_t xoff = 0, int64_t c_begin = 0, int64_t c_end = -1, int prod = 0,
                   const HaloFuse *hf = nullptr, int64_t batch = 1, int64_t batch_stride = 0) {
    if (n1_in < 0) n1_in = n1;
    if (c_end < 0) c_end = n1;
    if (c_end <= c_begin) { *rc = 0; return true; }
    constexpr int EPV = 16 / (int)sizeof(T);
    const int64_t nba = kla + kua + 1, nbb = klb + kub + 1;
    // instantiated band-count pairs: {1,3,5} x {1,3,5} (tri / pentadiagonal factors and their one-sided mode products) and
    // the 9-diagonal ((4,4)) family {9,9}, {9,1}, {1,9}
    const bool small = (nba == 3 || nba == 1 || nba == 5) && (nbb == 3 || nbb == 1 || nbb == 5);
    const bool nine = (nba == 9 && (nbb == 9 || nbb == 1)) || (nba == 1 && nbb == 9);
    const bool shape_ok = (small || nine) && !(nba == 1 && nbb == 1);
    if (batch > 65535 || (batch > 1 && (batch_stride % EPV != 0 || hf))) return false;
    if (!shape_ok || n2 % EPV != 0 || !lbm_aligned16(x) || n2 < 512 || (c_end < 0 ? n1 : c_end) - c_begin < 1 || h->tune.kron_force_tile) return false;
    KronRingArgs<T> a;
    a.n1 = n1; a.n2 = n2; a.kla = (int)kla; a.kua = (int)kua; a.klb = (int)klb; a.kub = (int)kub;
    a.lda = lda; a.ldb = ldb; a.A = A; a.B = B; a.alpha = alpha; a.beta = beta; a.x = x; a.y = y;
    a.n1_in = n1_in; a.xoff = xoff; a.c_begin = c_begin; a.c_end = c_end; a.prod = prod; a.batch_stride = batch_stride;
    const int64_t tr = kron_ring_tr(h);
    const int64_t hmax = klb > kub ? klb : kub;
    a.hp = (int)((hmax + EPV - 1) / EPV * EPV);
    a.tr = (int)tr;
    a.seg = (int)(tr + 2 * a.hp);
    if (nba == 3 && nbb == 3) *rc = kron_ring_launch2<T, 3, 3>(h, a, batch, hf);
    else if (nba == 1 && nbb == 3) *rc = kron_ring_launch2<T, 1, 3>(h, a, batch, hf);
    else if (nba == 3 && nbb == 1) *rc = kron_ring_launch2<T, 3, 1>(h, a, batch, hf);
    else if (nba == 5 && nbb == 5) *rc = kron_ring_launch2<T, 5, 5>(h, a, batch, hf);
    else if (nba == 3 && nbb == 5) *rc = kron_ring_launch2<T, 3, 5>(h, a, batch, hf);
    else if (nba == 5 && nbb == 3) *rc = kron_ring_launch2<T, 5, 3>(h, a, batch, hf);
    else if (nba == 5 && nbb == 1) *rc = kron_ring_launch2<T, 5, 1>(h, a, batch, hf);
    else if (nba == 1 && nbb == 5) *rc = kron_ring_launch2<T, 1, 5>(h, a, batch, hf);
    else if (nba == 9 && nbb == 9) *rc = kron_ring_launch2<T, 9, 9>(h, a, batch, hf);
    else if (nba == 9 && nbb == 1) *rc = kron_ring_launch2<T, 9, 1>(h, a, batch, hf);
    else *rc = kron_ring_launch2<T, 1, 9>(h, a, batch, hf);
    return true;
}

template <typename T>
int kron_launch(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const T *A, int64_t lda, int64_t klb,
                int64_t kub, const T *B, int64_t ldb, T alpha, const T *x, T beta, T *y, int useA, int useB,
                int64_t batch = 1, int64_t batch_stride = 0, int64_t n1_in = -1, int64_t xoff = 0, int64_t c_begin = 0,
                int64_t c_end = -1) {
    KronArgs<T> a;
    a.batch_stride = batch_stride;
    a.n1_in = n1_in < 0 ? n1 : n1_in;
    a.xoff = xoff;
    a.c_begin = c_begin;
    a.c_end = c_end < 0 ? n1 : c_end;
    if (a.c_end <= a.c_begin) return 0;
    a.n1 = n1; a.n2 = n2; a.kla = (int)kla; a.kua = (int)kua; a.klb = (int)klb; a.kub = (int)kub;
    a.lda = lda; a.ldb = ldb; a.A = A; a.B = B; a.alpha = alpha; a.beta = beta; a.x = x; a.y = y;
    a.useA = useA; a.useB = useB;
    int64_t tr = h->tune.kron_tr > 0 ? h->tune.kron_tr : 256;
    int64_t tc = h->tune.kron_tc > 0 ? h->tune.kron_tc : 16;
    if (tr > n2) tr = (n2 + 31) / 32 * 32;
    if (tc > n1) tc = n1;
    const int64_t smem_cap = 96 * 1024;
    auto layout = [&](int64_t tr_, int64_t tc_, int &pitch, int &oCA, int &oCB) {
        int64_t rows_h = tr_ + (useB ? klb + kub : 0);
        int64_t cols_h = tc_ + (useA ? kla + kua : 0);
        int64_t p = rows_h | 1;  // odd pitch: column-direction reads hit distinct banks
        pitch = (int)p;
        oCA = (int)(p * cols_h);
        oCB = (int)(oCA + tc_ * (kla + kua + 1));
        return (int64_t)(oCB + tr_ * (klb + kub + 1)) * (int64_t)sizeof(T);
    };
    int pitch, oCA, oCB;
    while (layout(tr, tc, pitch, oCA, oCB) > smem_cap && tc > 1) tc = (tc + 1) / 2;
    while (layout(tr, tc, pitch, oCA, oCB) > smem_cap && tr > 32) tr /= 2;
    const int64_t smem = layout(tr, tc, pitch, oCA, oCB);
    if (smem > smem_cap) return lbm_fail_arg(h, 4, "bands too wide for the fused Kronecker kernel");
    a.tr = (int)tr; a.tc = (int)tc; a.pitch = pitch; a.offCA = oCA; a.offCB = oCB;
    const int64_t gx = (n2 + tr - 1) / tr, gy = (a.c_end - a.c_begin + tc - 1) / tc;
    if (gy > 65535) return lbm_fail_arg(h, 2, "n1 too large for the tile grid");
    static std::atomic<int> configured_dev_mask{0};
    auto kern = kron_tile_kernel<T>;
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        configured_dev_mask |= (1 << h->device);
    }
    if (batch > 65535) return lbm_fail_arg(h, 2, "batch too large for the tile grid");
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)batch);
    kern<<<grid, 256, (size_t)smem, h->stream>>>(a);
    LBM_LAUNCH_CHECK(h, "kron_tile_kernel");
    return 0;
}

template <typename T>
int kron_check(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const T *A, int64_t lda, int64_t klb,
               int64_t kub, const T *B, int64_t ldb, const T *x, T *y) {
    if (!h) return -1;
    if (n1 < 0) return lbm_fail_arg(h, 2, "n1 < 0");
    if (n2 < 0) return lbm_fail_arg(h, 3, "n2 < 0");
    if (kla < 0) return lbm_fail_arg(h, 4, "kla < 0");
    if (kua < 0) return lbm_fail_arg(h, 5, "kua < 0");
    if (lda < kla + kua + 1) return lbm_fail_arg(h, 7, "lda < kla+kua+1");
    if (klb < 0) return lbm_fail_arg(h, 8, "klb < 0");
    if (kub < 0) return lbm_fail_arg(h, 9, "kub < 0");
    if (ldb < klb + kub + 1) return lbm_fail_arg(h, 11, "ldb < klb+kub+1");
    if (n1 == 0 || n2 == 0) return 0;
    if (!A) return lbm_fail_arg(h, 6, "A == NULL");
    if (!B) return lbm_fail_arg(h, 10, "B == NULL");
    if (!x) return lbm_fail_arg(h, 13, "x == NULL");
    if (!y) return lbm_fail_arg(h, 15, "y == NULL");
    return 0;
}

template <typename T>
int kronsum_impl(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const T *A, int64_t lda,
                 int64_t klb, int64_t kub, const T *B, int64_t ldb, T alpha, const T *x, T beta, T *y) {
    int rc = kron_check<T>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, x, y);
    if (rc || n1 == 0 || n2 == 0) return rc;
    lbm_device_guard g(h->device);
    if (kron_ring_try<T>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, x, beta, y, &rc)) return rc;
    return kron_launch<T>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, x, beta, y, 1, 1);
}

// Column-partitioned kron(A,I) + kron(I,B): rank r owns grid columns
// [r0,r1) of X (contiguous memory), stores the column slice [c0,c1) of A's band array, and keeps X
// in a ghost-cell buffer [kla ghost columns | owned | kua ghost columns].  The B*X term is rank
// local; only the ghost columns (n2 elements each) travel.  Interior columns overlap the exchange.
template <typename T>
int kronsum_range(lbm_handle_t h, int64_t m_loc, int64_t n_loc, int64_t left, int64_t n2, int64_t kla, int64_t kua,
                  const T *A, int64_t lda, int64_t klb, int64_t kub, const T *B, int64_t ldb, T alpha, const T *xbuf, T beta,
                  T *y, int64_t cb, int64_t ce, const HaloFuse *hf = nullptr) {
    if (ce <= cb) return hf ? LBM_NOT_FUSED : 0;
    int rc = 0;
    if (kron_ring_try<T>(h, m_loc, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, xbuf, beta, y, &rc, n_loc, left, cb, ce, 0, hf))
        return rc;
    if (hf) return LBM_NOT_FUSED;   // only the ring kernel carries the fused exchange; nothing was launched
    return kron_launch<T>(h, m_loc, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, xbuf, beta, y, 1, 1, 1, 0, n_loc, left,
                          cb, ce);
}

template <typename T>
int kronsum_sharded(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const T *A_slab, int64_t lda,
                    int64_t klb, int64_t kub, const T *B, int64_t ldb, T alpha, T *xbuf, T beta, T *y_own, int overlap) {
    int rc = kron_check<T>(h, n1, n2, kla, kua, A_slab, lda, klb, kub, B, ldb, xbuf, y_own);
    if (rc || n1 == 0 || n2 == 0) return rc;
    lbm_device_guard g(h->device);
    int64_t pl[10];
    HaloFuse hf;
    int fused = 0;
    // one flag word per row strip of the ring kernel (same `tr` rule as kron_ring_try; uniform across ranks); when the
    // strips do not fit the flag block everybody uses ONE word and the unfused exchange
    const int64_t tr = kron_ring_tr(h);
    const int64_t nstrips = (n2 + tr - 1) / tr;
    const int nflags = nstrips <= 4096 ? (int)nstrips : 1;
    if (nflags == 1 && nstrips > 1) overlap = overlap ? 2 : 0;   // 2: overlapped but never fused
    // plan + uniform validation + either a fused-exchange descriptor or the two-kernel / NCCL exchange already in flight
    if ((rc = lbm_internal_exchange(h, n1, kla, kua, xbuf, n2 * (int64_t)sizeof(T), pl, overlap == 1, overlap == 1 ? &hf : nullptr, &fused, nflags)))
        return rc;
    const int64_t m_loc = pl[9], n_loc = pl[3] - pl[2], left = pl[4], right = pl[5];
    if (m_loc == 0) return 0;
    if (h->world == 1 || pl[8] <= 1) {
        if (h->tune.exchange_fuse == 3) {   // self-test knob: the HALO instantiation with no neighbours (what the variant itself costs)
            HaloFuse none;
            memset(&none, 0, sizeof(none));
            none.on = 1 | (int)(h->tune.halo_timeout_ms & 6);   // self-test sub-knobs ride in halo_timeout_ms: 2 = no rotation, 4 = no patch test
            none.nflags = 1;
            rc = kronsum_range<T>(h, m_loc, n_loc, left, n2, kla, kua, A_slab, lda, klb, kub, B, ldb, alpha, xbuf, beta, y_own, 0, m_loc, &none);
            if (rc != LBM_NOT_FUSED) return rc;
        }
        return kronsum_range<T>(h, m_loc, n_loc, left, n2, kla, kua, A_slab, lda, klb, kub, B, ldb, alpha, xbuf, beta, y_own, 0, m_loc);
    }
    if (fused) {   // ONE launch: push at kernel start, ghost columns patched into the last chunks' stages
        rc = kronsum_range<T>(h, m_loc, n_loc, left, n2, kla, kua, A_slab, lda, klb, kub, B, ldb, alpha, xbuf, beta, y_own, 0, m_loc, &hf);
        if (rc != LBM_NOT_FUSED) return rc;
        if ((rc = lbm_internal_exchange_unfuse(h, n1, kla, kua, xbuf, n2 * (int64_t)sizeof(T), nflags))) return rc;
    }
    if (!overlap) {
        LBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
        return kronsum_range<T>(h, m_loc, n_loc, left, n2, kla, kua, A_slab, lda, klb, kub, B, ldb, alpha, xbuf, beta, y_own, 0, m_loc);
    }
    int64_t a = left == 0 ? 0 : kla, b = right == 0 ? m_loc : m_loc - kua;
    if (a > m_loc) a = m_loc;
    if (b < a) b = a;
    if ((rc = kronsum_range<T>(h, m_loc, n_loc, left, n2, kla, kua, A_slab, lda, klb, kub, B, ldb, alpha, xbuf, beta, y_own, a, b))) return rc;
    LBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
    if ((rc = kronsum_range<T>(h, m_loc, n_loc, left, n2, kla, kua, A_slab, lda, klb, kub, B, ldb, alpha, xbuf, beta, y_own, 0, a))) return rc;
    return kronsum_range<T>(h, m_loc, n_loc, left, n2, kla, kua, A_slab, lda, klb, kub, B, ldb, alpha, xbuf, beta, y_own, b, m_loc);
}

// y <- alpha*kron(A,B)*x + beta*y = alpha*vec(B*X*A^T) + beta*y  (BlockKron(A,B)*x; KronTrav(A,B)*DiagTrav(X) before the
// DiagTrav re-ordering, /root/reference/src/blockkron.jl:433-440)
template <typename T>
int kron2d_impl(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const T *A, int64_t lda, int64_t klb,
                int64_t kub, const T *B, int64_t ldb, T alpha, const T *x, T beta, T *y) {
    int rc = kron_check<T>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, x, y);
    if (rc || n1 == 0 || n2 == 0) return rc;
    lbm_device_guard g(h->device);
    // one pass when the streaming ring kernel covers the band shapes: y[r,c] = sum_tA A[c,.] * (sum_tB B[r,.] * X[., .]),
    // i.e. (B*X)*A^T entry by entry with X read once (2N elements of traffic instead of 4N + a scratch matrix)
    if (kron_ring_try<T>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, x, beta, y, &rc, -1, 0, 0, -1, 1)) return rc;
    constexpr int EPV = 16 / (int)sizeof(T);
    void *z;
    const int64_t nz = n1 > n2 ? n1 : n2;
    const int64_t zpad = (n1 * n2 + EPV - 1) / EPV * EPV;   // keeps the zero vector 16-byte aligned behind Z
    if ((rc = lbm_scratch(h, (size_t)(zpad + nz) * sizeof(T), &z))) return rc;
    T *Z = (T *)z, *zeros = Z + zpad;
    // Z = B*X, then Y = alpha*Z*A^T + beta*Y  (left to right, as src/blockkron.jl:440 evaluates B*X*A').
    // Each pass is a Kronecker SUM with a zero 1-band partner (X*0^T + B*X, then Z*A^T + 0*Z), which the streaming ring
    // kernel covers for tridiagonal / pentadiagonal factors; other shapes take the halo'd tile kernel.
    LBM_CUDA(h, cudaMemsetAsync(zeros, 0, (size_t)nz * sizeof(T), h->stream));
    if (!kron_ring_try<T>(h, n1, n2, 0, 0, zeros, 1, klb, kub, B, ldb, (T)1, x, (T)0, Z, &rc))
        rc = kron_launch<T>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, (T)1, x, (T)0, Z, 0, 1);
    if (rc) return rc;
    if (!kron_ring_try<T>(h, n1, n2, kla, kua, A, lda, 0, 0, zeros, 1, alpha, Z, beta, y, &rc))
        rc = kron_launch<T>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, Z, beta, y, 1, 0);
    return rc;
}

// Y = (A (x) B (x) C) applied to a full n x n x n tensor X (k fastest, then j, then l) as the three
// mode products of /root/reference/src/blockkron.jl:441-450: A along dim 3, B along dim 2, C along
// dim 1 -- the arithmetic of KronTrav(A,B,C)*DiagTrav(X) before the DiagTrav re-ordering.
template <typename T>
int kron3d_impl(lbm_handle_t h, int64_t n, int64_t kla, int64_t kua, const T *A, int64_t lda, int64_t klb, int64_t kub,
                const T *B, int64_t ldb, int64_t klc, int64_t kuc, const T *C, int64_t ldc, T alpha, const T *x, T beta, T *y) {
    if (!h) return -1;
    if (n < 0) return lbm_fail_arg(h, 2, "n < 0");
    if (kla < 0 || kua < 0 || lda < kla + kua + 1) return lbm_fail_arg(h, 6, "bad band of A");
    if (klb < 0 || kub < 0 || ldb < klb + kub + 1) return lbm_fail_arg(h, 10, "bad band of B");
    if (klc < 0 || kuc < 0 || ldc < klc + kuc + 1) return lbm_fail_arg(h, 14, "bad band of C");
    if (n == 0) return 0;
    if (!A) return lbm_fail_arg(h, 5, "A == NULL");
    if (!B) return lbm_fail_arg(h, 9, "B == NULL");
    if (!C) return lbm_fail_arg(h, 13, "C == NULL");
    if (!x) return lbm_fail_arg(h, 16, "x == NULL");
    if (!y) return lbm_fail_arg(h, 18, "y == NULL");
    lbm_device_guard g(h->device);
    constexpr int EPV = 16 / (int)sizeof(T);
    void *z1v;
    int rc;
    const int64_t n3 = n * n * n, n3p = (n3 + EPV - 1) / EPV * EPV, nz = n * n;
    if ((rc = lbm_scratch(h, (size_t)(2 * n3p + nz + EPV) * sizeof(T), &z1v))) return rc;
    T *z1 = (T *)z1v, *z2 = z1 + n3p, *zeros = z2 + n3p;
    LBM_CUDA(h, cudaMemsetAsync(zeros, 0, (size_t)nz * sizeof(T), h->stream));
    // mode 3: X as an (n^2 x n) matrix, Z1 = X * A^T -- a Kronecker sum with a zero 1-band partner on the long side, which
    // the streaming ring kernel takes in ONE launch when A is tri/pentadiagonal (n^2 rows per grid column)
    if (!kron_ring_try<T>(h, n, n * n, kla, kua, A, lda, 0, 0, zeros, 1, (T)1, x, (T)0, z1, &rc))
        rc = kron_launch<T>(h, n, n * n, kla, kua, A, lda, 0, 0, A, lda, (T)1, x, (T)0, z1, 1, 0);
    if (rc) return rc;
    // mode 2: each l-slice (n x n): Z2[:,:,l] = Z1[:,:,l] * B^T   (batched over l: blockIdx.z of the ring kernel)
    if (!kron_ring_try<T>(h, n, n, klb, kub, B, ldb, 0, 0, zeros, 1, (T)1, z1, (T)0, z2, &rc, -1, 0, 0, -1, 0, nullptr, n, n * n))
        rc = kron_launch<T>(h, n, n, klb, kub, B, ldb, 0, 0, B, ldb, (T)1, z1, (T)0, z2, 1, 0, n, n * n);
    if (rc) return rc;
    // mode 1: Z2 as an (n x n^2) matrix, Y = alpha * C * Z2 + beta * Y: the ring kernel again (X*0^T + C*X over n^2 columns)
    if (!kron_ring_try<T>(h, n * n, n, 0, 0, zeros, 1, klc, kuc, C, ldc, alpha, z2, beta, y, &rc))
        rc = kron_launch<T>(h, n * n, n, 0, 0, C, ldc, klc, kuc, C, ldc, alpha, z2, beta, y, 0, 1);
    return rc;
}

}  // namespace

extern "C" {
int lbm_dkronsum2d_mv_sharded(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const double *A_slab,
                              int64_t lda, int64_t klb, int64_t kub, const double *B, int64_t ldb, double alpha,
                              double *xbuf, double beta, double *y_own, int overlap) {
    return kronsum_sharded<double>(h, n1, n2, kla, kua, A_slab, lda, klb, kub, B, ldb, alpha, xbuf, beta, y_own, overlap);
}
int lbm_skronsum2d_mv_sharded(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const float *A_slab,
                              int64_t lda, int64_t klb, int64_t kub, const float *B, int64_t ldb, float alpha, float *xbuf,
                              float beta, float *y_own, int overlap) {
    return kronsum_sharded<float>(h, n1, n2, kla, kua, A_slab, lda, klb, kub, B, ldb, alpha, xbuf, beta, y_own, overlap);
}
int lbm_dkronsum2d_mv(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const double *A, int64_t lda,
                      int64_t klb, int64_t kub, const double *B, int64_t ldb, double alpha, const double *x,
                      double beta, double *y) {
    return kronsum_impl<double>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, x, beta, y);
}
int lbm_skronsum2d_mv(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const float *A, int64_t lda,
                      int64_t klb, int64_t kub, const float *B, int64_t ldb, float alpha, const float *x, float beta,
                      float *y) {
    return kronsum_impl<float>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, x, beta, y);
}
int lbm_dkron2d_mv(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const double *A, int64_t lda,
                   int64_t klb, int64_t kub, const double *B, int64_t ldb, double alpha, const double *x, double beta,
                   double *y) {
    return kron2d_impl<double>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, x, beta, y);
}
int lbm_skron2d_mv(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const float *A, int64_t lda,
                   int64_t klb, int64_t kub, const float *B, int64_t ldb, float alpha, const float *x, float beta, float *y) {
    return kron2d_impl<float>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, x, beta, y);
}
int lbm_dkron3d_mv(lbm_handle_t h, int64_t n, int64_t kla, int64_t kua, const double *A, int64_t lda, int64_t klb,
                   int64_t kub, const double *B, int64_t ldb, int64_t klc, int64_t kuc, const double *C, int64_t ldc,
                   double alpha, const double *x, double beta, double *y) {
    return kron3d_impl<double>(h, n, kla, kua, A, lda, klb, kub, B, ldb, klc, kuc, C, ldc, alpha, x, beta, y);
}
int lbm_skron3d_mv(lbm_handle_t h, int64_t n, int64_t kla, int64_t kua, const float *A, int64_t lda, int64_t klb,
                   int64_t kub, const float *B, int64_t ldb, int64_t klc, int64_t kuc, const float *C, int64_t ldc,
                   float alpha, const float *x, float beta, float *y) {
    return kron3d_impl<float>(h, n, kla, kua, A, lda, klb, kub, B, ldb, klc, kuc, C, ldc, alpha, x, beta, y);
}
}
