// gbmv.cu -- y <- alpha*op(A)*x + beta*y for a banded A in BandedMatrices' diagonal-major
// (kl+ku+1) x n storage.  Replaces BandedMatrices.gbmv! -> BLAS ?gbmv_,
// the leaf of every `mul!` in the reference's hot path.
//
// HBM-bound (0.25 flop/B): every kernel here is organised around moving the band data
// exactly once at full bandwidth.
//
//  * gbmv_tile_kernel   (narrow bands, kl+ku+1 <= ~33): persistent CTAs walk output tiles.
//    The tile's slab of A -- columns [o0-kl, o1+ku), ONE contiguous byte range of the
//    diagonal-major array -- and its x window are fetched by 1-D bulk TMA (UBLKCP) into a
//    multi-stage shared-memory ring guarded by mbarriers; threads then read one data element
//    per FMA from smem (stride lda between lanes: conflict-free for odd lda) and write y
//    coalesced.  No thread ever issues a global load for A.
//  * gbmv_wide_n_kernel (wide bands, op = N): a CTA owns a contiguous run of rows and streams
//    the columns it needs through the smem ring in chunks of C columns; row i lives in thread
//    (i - R0) mod NT with its accumulator in a register for the ~(kl+ku+C)/C chunks that
//    touch it, so each element of A is read from smem by exactly one thread, lanes read
//    consecutive addresses, and completed rows retire as coalesced stores.
//  * gbmv_wide_t_ring_kernel (wide bands, op = T): column chunks (contiguous) + x window through
//    the same TMA ring; one warp reduces a whole column from smem with a shuffle tree.
//  * gbmv_wide_t_kernel: unaligned fallback of the above, reading global memory directly.
//  * gbmv_direct_kernel: any alignment / any lda fallback, thread per output, L1/L2 served.
#include "halo.cuh"

namespace {

using namespace lbm;

constexpr int GBMV_NT = 256;
constexpr int GBMV_MAX_STAGES = 8;

template <typename T>
struct GbmvArgs {
    int64_t m, n, kl, ku, lda;
    int64_t nout;  // outputs: m (N) or n (T)
    T alpha, beta;
    const T *A;
    const T *x;
    T *y;
    int64_t tile;    // outputs per tile (multiple of GBMV_NT)
    int64_t ntiles;
    int stages;
    int64_t sA_elems, sx_elems;  // per-stage smem capacity in elements (multiples of 16B)
    int64_t skew;    // experiment: start every bulk copy this many elements (a multiple of 16 B) before its natural aligned start
};

// ------------------------------------------------------------------------------------------
// narrow bands: streamed tiles
// ------------------------------------------------------------------------------------------
// HALO: the row-sharded apply in ONE launch (halo.cuh) -- CTA 0 pushes this rank's boundary values of x to the
// neighbours' mailboxes first; tiles are walked rotated by one so that the tiles whose x window overlaps ghost cells
// (the first and the last) come last, and those tiles patch their window in shared memory from the local mailbox.
template <typename T, int NB, bool TRANS, bool HALO>
__global__ void __launch_bounds__(GBMV_NT) gbmv_tile_kernel(const GbmvArgs<T> a, const __grid_constant__ HaloFuse hf) {
    constexpr int EPV = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    T *sbase = reinterpret_cast<T *>(smem_raw + 128);
    const int tid = threadIdx.x;
    const int S = a.stages;
    const int nb = NB > 0 ? NB : (int)(a.kl + a.ku + 1);
    const int64_t stage_elems = a.sA_elems + a.sx_elems;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    pdl_launch_dependents();   // the next launch in the stream may start placing its CTAs as ours retire ...
    __syncthreads();
    pdl_wait();                // ... and we touch no global memory before our own predecessor has completed

    // tiles of this CTA: t = blockIdx.x + k*gridDim.x
    const int64_t first = blockIdx.x;
    const int64_t step = gridDim.x;
    const int64_t mine = first < a.ntiles ? (a.ntiles - first + step - 1) / step : 0;

    auto ranges = [&](int64_t t, int64_t &o0, int64_t &o1, int64_t &c0, int64_t &c1, int64_t &x0, int64_t &x1) {
        if (HALO && !(hf.on & 2)) t = t + 1 < a.ntiles ? t + 1 : 0;   // rotate: tile 0 (left ghosts) last, tile ntiles-1 second to last
        o0 = t * a.tile;
        o1 = o0 + a.tile < a.nout ? o0 + a.tile : a.nout;
        if (!TRANS) {
            c0 = o0 - a.kl > 0 ? o0 - a.kl : 0;
            c1 = o1 + a.ku < a.n ? o1 + a.ku : a.n;
            if (c0 > a.n) c0 = a.n;
            if (c1 < c0) c1 = c0;
            x0 = c0;
            x1 = c1;
        } else {
            c0 = o0;
            c1 = o1;
            x0 = o0 - a.ku > 0 ? o0 - a.ku : 0;
            x1 = o1 + a.kl < a.m ? o1 + a.kl : a.m;
            if (x0 > a.m) x0 = a.m;
            if (x1 < x0) x1 = x0;
        }
    };

    auto issue = [&](int64_t t, int s) {  // thread 0 only
        int64_t o0, o1, c0, c1, x0, x1;
        ranges(t, o0, o1, c0, c1, x0, x1);
        int64_t e0a = (c0 * a.lda) & ~(int64_t)(EPV - 1);
        int64_t x0a = x0 & ~(int64_t)(EPV - 1);
        if (a.skew > 0 && e0a >= a.skew) e0a -= a.skew;
        if (a.skew > 0 && x0a >= a.skew) x0a -= a.skew;
        const int64_t cntA = c1 * a.lda - e0a;
        const int64_t cntx = x1 - x0a;
        T *sA = sbase + (int64_t)s * stage_elems;
        T *sx = sA + a.sA_elems;
        mbar_add_tx(&bars[s], stage_bytes_bulk<T>(cntA) + stage_bytes_bulk<T>(cntx));
        stage_issue<T>(sA, a.A + e0a, cntA, &bars[s]);
        stage_issue<T>(sx, a.x + x0a, cntx, &bars[s]);
        if (a.skew < 0) __nanosleep((unsigned)(-a.skew));   // experiment: pace the issuing thread (gbmv_skew < 0 = ns)
        mbar_arrive(&bars[s]);
    };

    if (tid == 0) {
        const int64_t pro = mine < S ? mine : S;
        for (int64_t k = 0; k < pro; ++k) issue(first + k * step, (int)k);
    }
    // the push rides on the latency of the prologue copies already in flight (the neighbours need it only for their LAST tiles)
    if (HALO && blockIdx.x == 0) halo_push(hf, tid, GBMV_NT);

    const int lda = (int)a.lda;
    for (int64_t k = 0; k < mine; ++k) {
        const int s = (int)(k % S);
        const uint32_t parity = (uint32_t)((k / S) & 1);
        const int64_t t = first + k * step;
        int64_t o0, o1, c0, c1, x0, x1;
        ranges(t, o0, o1, c0, c1, x0, x1);
        int64_t e0a = (c0 * a.lda) & ~(int64_t)(EPV - 1);
        int64_t x0a = x0 & ~(int64_t)(EPV - 1);
        if (a.skew > 0 && e0a >= a.skew) e0a -= a.skew;
        if (a.skew > 0 && x0a >= a.skew) x0a -= a.skew;
        const T *sA = sbase + (int64_t)s * stage_elems;
        const T *sx = sA + a.sA_elems;

        mbar_wait(&bars[s], parity);
        // ghost cells can only sit in the x windows of the first and of the last two tiles (halo <= 128 < tile)
        if (HALO && !(hf.on & 4) && (o0 == 0 || o1 + 2 * a.tile > a.nout)) halo_patch_1d<T>(hf, const_cast<T *>(sx), x0a, x1, tid, GBMV_NT);

        // interior tiles need no index masks
        const bool interior = TRANS ? (o0 - a.ku >= 0 && o1 - 1 + a.kl < a.m)
                                    : (o0 - a.kl >= 0 && o1 - 1 + a.ku < a.n);
        for (int64_t o = o0 + tid; o < o1; o += GBMV_NT) {
            T acc = 0;
            if (!TRANS) {
                // terms j = o-kl .. o+ku (ascending), band row r = ku + o - j
                const int jb = (int)(o - a.kl - x0a);                         // sx index of j = o-kl
                const int ab = (int)((a.kl + a.ku) + (o - a.kl) * a.lda - e0a);  // sA index of (r=kl+ku, j=o-kl)
                if (interior) {
#pragma unroll
                    for (int q = 0; q < nb; ++q) acc = fma(sA[ab + q * (lda - 1)], sx[jb + q], acc);
                } else {
#pragma unroll
                    for (int q = 0; q < nb; ++q) {
                        const int64_t j = o - a.kl + q;
                        if (j >= 0 && j < a.n) acc = fma(sA[ab + q * (lda - 1)], sx[jb + q], acc);
                    }
                }
            } else {
                // terms i = o-ku .. o+kl (ascending), band row r = ku + i - o
                const int ib = (int)(o - a.ku - x0a);
                const int ab = (int)(o * a.lda - e0a);
                if (interior) {
#pragma unroll
                    for (int q = 0; q < nb; ++q) acc = fma(sA[ab + q], sx[ib + q], acc);
                } else {
#pragma unroll
                    for (int q = 0; q < nb; ++q) {
                        const int64_t i = o - a.ku + q;
                        if (i >= 0 && i < a.m) acc = fma(sA[ab + q], sx[ib + q], acc);
                    }
                }
            }
            a.y[o] = (a.beta == (T)0) ? a.alpha * acc : fma(a.beta, a.y[o], a.alpha * acc);
        }
        __syncthreads();  // every thread is done reading stage s
        if (tid == 0 && k + S < mine) issue(first + (k + S) * step, s);
    }
    if (HALO && blockIdx.x == 0) halo_final_wait(hf, tid);
}

// ------------------------------------------------------------------------------------------
// wide bands, op = N: column-streaming with register-resident row accumulators
// ------------------------------------------------------------------------------------------
template <typename T>
struct GbmvWideArgs {
    int64_t m, n, kl, ku, lda;
    T alpha, beta;
    const T *A;
    const T *x;
    T *y;
    int64_t run;    // rows per CTA
    int chunk;      // columns per chunk (multiple of 16B worth of elements)
    int stages;
    int64_t sA_elems, sx_elems;
};

template <typename T>
__global__ void __launch_bounds__(1024) gbmv_wide_n_kernel(const GbmvWideArgs<T> a) {
    constexpr int EPV = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    T *sbase = reinterpret_cast<T *>(smem_raw + 128);
    const int tid = threadIdx.x;
    const int NT = blockDim.x;
    const int S = a.stages;
    const int64_t stage_elems = a.sA_elems + a.sx_elems;
    const int C = a.chunk;

    const int64_t R0 = (int64_t)blockIdx.x * a.run;
    const int64_t R1 = R0 + a.run < a.m ? R0 + a.run : a.m;
    if (R0 >= R1) return;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    int64_t cs = R0 - a.kl > 0 ? R0 - a.kl : 0;
    cs &= ~(int64_t)(EPV - 1);
    int64_t ce = R1 + a.ku < a.n ? R1 + a.ku : a.n;  // exclusive
    if (cs > ce) cs = ce & ~(int64_t)(EPV - 1);
    const int64_t nchunks = (ce - cs + C - 1) / C;

    auto issue = [&](int64_t q, int s) {  // thread 0 only
        const int64_t cb = cs + q * C;
        const int64_t cq = cb + C < ce ? C : ce - cb;
        T *sA = sbase + (int64_t)s * stage_elems;
        T *sx = sA + a.sA_elems;
        const int64_t cntA = cq * a.lda;
        mbar_add_tx(&bars[s], stage_bytes_bulk<T>(cntA) + stage_bytes_bulk<T>(cq));
        stage_issue<T>(sA, a.A + cb * a.lda, cntA, &bars[s]);
        stage_issue<T>(sx, a.x + cb, cq, &bars[s]);
        mbar_arrive(&bars[s]);
    };

    if (tid == 0) {
        const int64_t pro = nchunks < S ? nchunks : S;
        for (int64_t q = 0; q < pro; ++q) issue(q, (int)q);
    }

    int64_t i = R0 + tid;  // the row this thread currently accumulates
    T acc = 0;
    const int lda = (int)a.lda;

    for (int64_t q = 0; q < nchunks; ++q) {
        const int s = (int)(q % S);
        const uint32_t parity = (uint32_t)((q / S) & 1);
        const int64_t cb = cs + q * C;
        const int cq = (int)(cb + C < ce ? C : ce - cb);
        const T *sA = sbase + (int64_t)s * stage_elems;
        const T *sx = sA + a.sA_elems;

        mbar_wait(&bars[s], parity);

        if (i < R1) {
            // columns of this chunk that touch row i: j in [i-kl, i+ku]
            int jlo = (int)(i - a.kl - cb);
            int jhi = (int)(i + a.ku - cb);
            if (jlo < 0) jlo = 0;
            if (jhi > cq - 1) jhi = cq - 1;
            // sA index of (row i, column cb+jj): (ku + i - cb - jj) + lda*jj = base + jj*(lda-1)
            const int base = (int)(a.ku + i - cb);
#pragma unroll 4
            for (int jj = jlo; jj <= jhi; ++jj) acc = fma(sA[base + jj * (lda - 1)], sx[jj], acc);
            if (i + a.ku < cb + cq || q == nchunks - 1) {  // all columns of row i seen
                a.y[i] = (a.beta == (T)0) ? a.alpha * acc : fma(a.beta, a.y[i], a.alpha * acc);
                i += NT;
                acc = 0;
            }
        }
        __syncthreads();
        if (tid == 0 && q + S < nchunks) issue(q + S, s);
    }
    // rows with an empty column range (m > n + kl) or an empty sweep
    for (; i < R1; i += NT) a.y[i] = (a.beta == (T)0) ? (T)0 : a.beta * a.y[i];
}

// Wide bands, op = N, R right-hand sides sharing ONE pass over A (mul!(C, A, B::Matrix) with l+u+1 > 33): the kernel
// above with R accumulators per row and R x-windows per stage.  A is the whole traffic here (R*2/(l+u+1) of it is
// x and y), so R columns cost about one gbmv instead of R.
template <typename T> struct vec2w;
template <> struct vec2w<double> { using type = double2; };
template <> struct vec2w<float> { using type = float2; };

template <typename T>
struct GbmvWideMultiArgs {
    GbmvWideArgs<T> w;   // x / y = first column of X / Y
    int64_t ldx, ldy;
};

// MAXT = 512 when the CTA has <= 512 threads: 128 registers per thread instead of 64, i.e. room for the unrolled loads
template <typename T, int R, int MAXT>
__global__ void __launch_bounds__(MAXT) gbmv_wide_n_multi_kernel(const GbmvWideMultiArgs<T> am) {
    const GbmvWideArgs<T> &a = am.w;
    constexpr int EPV = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    T *sbase = reinterpret_cast<T *>(smem_raw + 128);
    const int tid = threadIdx.x;
    const int NT = blockDim.x;
    const int S = a.stages;
    const int sxe = (int)a.sx_elems;
    const int64_t stage_elems = a.sA_elems + R * a.sx_elems;
    const int C = a.chunk;

    const int64_t R0 = (int64_t)blockIdx.x * a.run;
    const int64_t R1 = R0 + a.run < a.m ? R0 + a.run : a.m;
    if (R0 >= R1) return;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    int64_t cs = R0 - a.kl > 0 ? R0 - a.kl : 0;
    cs &= ~(int64_t)(EPV - 1);
    int64_t ce = R1 + a.ku < a.n ? R1 + a.ku : a.n;  // exclusive
    if (cs > ce) cs = ce & ~(int64_t)(EPV - 1);
    const int64_t nchunks = (ce - cs + C - 1) / C;

    auto issue = [&](int64_t q, int s) {  // thread 0 only
        const int64_t cb = cs + q * C;
        const int64_t cq = cb + C < ce ? C : ce - cb;
        T *sA = sbase + (int64_t)s * stage_elems;
        T *sx = sA + a.sA_elems;
        const int64_t cntA = stage_round<T>(cq * a.lda, (a.n - cb) * a.lda);
        const int64_t cntx = stage_round<T>(cq, a.n - cb);
        mbar_add_tx(&bars[s], stage_bytes_bulk<T>(cntA) + R * stage_bytes_bulk<T>(cntx));
        stage_issue<T>(sA, a.A + cb * a.lda, cntA, &bars[s]);
#pragma unroll
        for (int r = 0; r < R; ++r) stage_issue<T>(sx + r * a.sx_elems, a.x + r * am.ldx + cb, cntx, &bars[s]);
        mbar_arrive(&bars[s]);
    };

    if (tid == 0) {
        const int64_t pro = nchunks < S ? nchunks : S;
        for (int64_t q = 0; q < pro; ++q) issue(q, (int)q);
    }

    int64_t i = R0 + tid;  // the row this thread currently accumulates
    T acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = (T)0;
    const int lda = (int)a.lda;

    for (int64_t q = 0; q < nchunks; ++q) {
        const int s = (int)(q % S);
        const uint32_t parity = (uint32_t)((q / S) & 1);
        const int64_t cb = cs + q * C;
        const int cq = (int)(cb + C < ce ? C : ce - cb);
        const T *sA = sbase + (int64_t)s * stage_elems;
        const T *sx = sA + a.sA_elems;

        mbar_wait(&bars[s], parity);

        if (i < R1) {
            int jlo = (int)(i - a.kl - cb);
            int jhi = (int)(i + a.ku - cb);
            if (jlo < 0) jlo = 0;
            if (jhi > cq - 1) jhi = cq - 1;
            const int base = (int)(a.ku + i - cb);
            // two columns per step: their x values are one aligned 2-element broadcast load per right-hand side
            using V2 = typename vec2w<T>::type;
            int jj = jlo;
            if ((jj & 1) && jj <= jhi) {
                const T av = sA[base + jj * (lda - 1)];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fma(av, sx[r * sxe + jj], acc[r]);
                ++jj;
            }
#pragma unroll 4
            for (; jj + 1 <= jhi; jj += 2) {
                const T av0 = sA[base + jj * (lda - 1)], av1 = sA[base + (jj + 1) * (lda - 1)];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const V2 xv = *reinterpret_cast<const V2 *>(sx + r * sxe + jj);
                    acc[r] = fma(av0, xv.x, acc[r]);
                    acc[r] = fma(av1, xv.y, acc[r]);
                }
            }
            if (jj <= jhi) {
                const T av = sA[base + jj * (lda - 1)];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fma(av, sx[r * sxe + jj], acc[r]);
            }
            if (i + a.ku < cb + cq || q == nchunks - 1) {  // all columns of row i seen
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    T *yp = a.y + r * am.ldy + i;
                    *yp = (a.beta == (T)0) ? a.alpha * acc[r] : fma(a.beta, *yp, a.alpha * acc[r]);
                    acc[r] = (T)0;
                }
                i += NT;
            }
        }
        __syncthreads();
        if (tid == 0 && q + S < nchunks) issue(q + S, s);
    }
    for (; i < R1; i += NT)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            T *yp = a.y + r * am.ldy + i;
            *yp = (a.beta == (T)0) ? (T)0 : a.beta * *yp;
        }
}

// ------------------------------------------------------------------------------------------
// wide bands, op = T, aligned: column chunks through the TMA ring, warp per column from smem
// ------------------------------------------------------------------------------------------
// y[j] = sum_r A[r, j] * x[j + r - ku].  A CTA owns a contiguous run of output columns; chunks of
// C columns (contiguous in the diagonal-major array) and the matching x window
// [cb - ku, cb + C + kl) arrive by bulk TMA; lane <-> column, warp <-> slice of the band rows
// (one smem read of A per FMA, conflict-free for odd lda), partial sums combined per chunk.
template <typename T>
__global__ void __launch_bounds__(256) gbmv_wide_t_ring_kernel(const GbmvWideArgs<T> a) {
    constexpr int EPV = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    T *sbase = reinterpret_cast<T *>(smem_raw + 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.stages;
    const int64_t stage_elems = a.sA_elems + a.sx_elems;
    const int C = a.chunk;
    const int nb = (int)(a.kl + a.ku + 1);

    const int64_t J0 = (int64_t)blockIdx.x * a.run;            // run is a multiple of EPV
    const int64_t J1 = J0 + a.run < a.n ? J0 + a.run : a.n;
    if (J0 >= J1) return;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t nchunks = (J1 - J0 + C - 1) / C;

    auto issue = [&](int64_t q, int s) {  // thread 0 only
        const int64_t cb = J0 + q * C;
        const int64_t cq = cb + C < J1 ? C : J1 - cb;
        int64_t x0 = cb - a.ku > 0 ? cb - a.ku : 0;
        x0 &= ~(int64_t)(EPV - 1);
        int64_t x1 = cb + cq + a.kl < a.m ? cb + cq + a.kl : a.m;
        if (x1 < x0) x1 = x0;
        T *sA = sbase + (int64_t)s * stage_elems;
        T *sx = sA + a.sA_elems;
        mbar_add_tx(&bars[s], stage_bytes_bulk<T>(cq * a.lda) + stage_bytes_bulk<T>(x1 - x0));
        stage_issue<T>(sA, a.A + cb * a.lda, cq * a.lda, &bars[s]);
        stage_issue<T>(sx, a.x + x0, x1 - x0, &bars[s]);
        mbar_arrive(&bars[s]);
    };
    if (tid == 0) {
        const int64_t pro = nchunks < S ? nchunks : S;
        for (int64_t q = 0; q < pro; ++q) issue(q, (int)q);
    }
    const int lda = (int)a.lda;
    const int rw = (nb + 7) / 8;                             // band rows per warp
    T *spart = sbase + (int64_t)S * stage_elems;             // [8][C] partial sums
    for (int64_t q = 0; q < nchunks; ++q) {
        const int s = (int)(q % S);
        const uint32_t parity = (uint32_t)((q / S) & 1);
        const int64_t cb = J0 + q * C;
        const int cq = (int)(cb + C < J1 ? C : J1 - cb);
        int64_t x0 = cb - a.ku > 0 ? cb - a.ku : 0;
        x0 &= ~(int64_t)(EPV - 1);
        const T *sA = sbase + (int64_t)s * stage_elems;
        const T *sx = sA + a.sA_elems;
        mbar_wait(&bars[s], parity);
        // lane <-> column (stride lda between lanes: conflict-free for odd lda), warp <-> a slice
        // of the band rows; the 8 partial sums per column meet in smem once per chunk.
        for (int jj = lane; jj < cq; jj += 32) {
            const int64_t j = cb + jj;
            const T *col = sA + jj * lda;
            const int xoff = (int)(j - a.ku - x0);            // sx index of row i = j - ku (r = 0)
            int rlo = (int)(a.ku - j > 0 ? a.ku - j : 0);     // valid band rows: 0 <= j + r - ku < m
            int rhi = (int)(a.m - 1 - j + a.ku < nb - 1 ? a.m - 1 - j + a.ku : nb - 1);
            const int w0 = warp * rw, w1 = w0 + rw - 1;
            if (rlo < w0) rlo = w0;
            if (rhi > w1) rhi = w1;
            T acc0 = 0, acc1 = 0;
            int r = rlo;
            for (; r + 1 <= rhi; r += 2) {
                acc0 = fma(col[r], sx[xoff + r], acc0);
                acc1 = fma(col[r + 1], sx[xoff + r + 1], acc1);
            }
            if (r <= rhi) acc0 = fma(col[r], sx[xoff + r], acc0);
            spart[warp * C + jj] = acc0 + acc1;
        }
        __syncthreads();
        if (tid < cq) {
            T acc = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) acc += spart[w * C + tid];
            const int64_t j = cb + tid;
            a.y[j] = (a.beta == (T)0) ? a.alpha * acc : fma(a.beta, a.y[j], a.alpha * acc);
        }
        __syncthreads();
        if (tid == 0 && q + S < nchunks) issue(q + S, s);
    }
}

// The same for R right-hand sides (mul!(C, A', B::Matrix) with wide bands; round 1 ran it column by column: R passes over A).
// One slab of A and R x-windows per stage; a lane keeps R accumulators for its column, so every value of A read from shared
// memory feeds R FMAs ((R+1)/R shared-memory reads per FMA instead of 2) and A crosses HBM once per R columns.
template <typename T, int R>
__global__ void __launch_bounds__(256) gbmv_wide_t_multi_kernel(const GbmvWideMultiArgs<T> am) {
    const GbmvWideArgs<T> &a = am.w;
    constexpr int EPV = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    T *sbase = reinterpret_cast<T *>(smem_raw + 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.stages;
    const int sxe = (int)a.sx_elems;
    const int64_t stage_elems = a.sA_elems + R * a.sx_elems;
    const int C = a.chunk;
    const int nb = (int)(a.kl + a.ku + 1);

    const int64_t J0 = (int64_t)blockIdx.x * a.run;            // run is a multiple of EPV
    const int64_t J1 = J0 + a.run < a.n ? J0 + a.run : a.n;
    if (J0 >= J1) return;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t nchunks = (J1 - J0 + C - 1) / C;

    auto issue = [&](int64_t q, int s) {  // thread 0 only
        const int64_t cb = J0 + q * C;
        const int64_t cq = cb + C < J1 ? C : J1 - cb;
        int64_t x0 = cb - a.ku > 0 ? cb - a.ku : 0;
        x0 &= ~(int64_t)(EPV - 1);
        int64_t x1 = cb + cq + a.kl < a.m ? cb + cq + a.kl : a.m;
        if (x1 < x0) x1 = x0;
        T *sA = sbase + (int64_t)s * stage_elems;
        T *sx = sA + a.sA_elems;
        const int64_t cntA = stage_round<T>(cq * a.lda, (a.n - cb) * a.lda);
        const int64_t cntx = stage_round<T>(x1 - x0, a.m - x0);
        mbar_add_tx(&bars[s], stage_bytes_bulk<T>(cntA) + R * stage_bytes_bulk<T>(cntx));
        stage_issue<T>(sA, a.A + cb * a.lda, cntA, &bars[s]);
#pragma unroll
        for (int r = 0; r < R; ++r) stage_issue<T>(sx + r * a.sx_elems, a.x + r * am.ldx + x0, cntx, &bars[s]);
        mbar_arrive(&bars[s]);
    };
    if (tid == 0) {
        const int64_t pro = nchunks < S ? nchunks : S;
        for (int64_t q = 0; q < pro; ++q) issue(q, (int)q);
    }
    const int lda = (int)a.lda;
    const int rw = (nb + 7) / 8;                             // band rows per warp
    T *spart = sbase + (int64_t)S * stage_elems;             // [8][R][C] partial sums
    for (int64_t q = 0; q < nchunks; ++q) {
        const int s = (int)(q % S);
        const uint32_t parity = (uint32_t)((q / S) & 1);
        const int64_t cb = J0 + q * C;
        const int cq = (int)(cb + C < J1 ? C : J1 - cb);
        int64_t x0 = cb - a.ku > 0 ? cb - a.ku : 0;
        x0 &= ~(int64_t)(EPV - 1);
        const T *sA = sbase + (int64_t)s * stage_elems;
        const T *sx = sA + a.sA_elems;
        mbar_wait(&bars[s], parity);
        for (int jj = lane; jj < cq; jj += 32) {
            const int64_t j = cb + jj;
            const T *col = sA + jj * lda;
            const int xoff = (int)(j - a.ku - x0);            // sx index of row i = j - ku (r = 0)
            int rlo = (int)(a.ku - j > 0 ? a.ku - j : 0);     // valid band rows: 0 <= j + r - ku < m
            int rhi = (int)(a.m - 1 - j + a.ku < nb - 1 ? a.m - 1 - j + a.ku : nb - 1);
            const int w0 = warp * rw, w1 = w0 + rw - 1;
            if (rlo < w0) rlo = w0;
            if (rhi > w1) rhi = w1;
            T acc[R];
#pragma unroll
            for (int c = 0; c < R; ++c) acc[c] = (T)0;
#pragma unroll 2
            for (int r = rlo; r <= rhi; ++r) {
                const T av = col[r];
                const T *xr = sx + xoff + r;
#pragma unroll
                for (int c = 0; c < R; ++c) acc[c] = fma(av, xr[c * sxe], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < R; ++c) spart[(warp * R + c) * C + jj] = acc[c];
        }
        __syncthreads();
        for (int e = tid; e < R * cq; e += 256) {
            const int c = e / cq, jj = e - c * cq;
            T acc = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) acc += spart[(w * R + c) * C + jj];
            T *yp = a.y + c * am.ldy + cb + jj;
            *yp = (a.beta == (T)0) ? a.alpha * acc : fma(a.beta, *yp, a.alpha * acc);
        }
        __syncthreads();
        if (tid == 0 && q + S < nchunks) issue(q + S, s);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) gbmv_wide_t_kernel(int64_t m, int64_t n, int64_t kl, int64_t ku, T alpha,
                                                         const T *__restrict__ A, int64_t lda,
                                                         const T *__restrict__ x, T beta, T *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nb = (int)(kl + ku + 1);
    for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += warps) {
        const T *col = A + j * lda;
        T acc = 0;
        for (int r = lane; r < nb; r += 32) {
            const int64_t i = j + r - ku;
            if (i >= 0 && i < m) acc = fma(col[r], x[i], acc);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) y[j] = (beta == (T)0) ? alpha * acc : fma(beta, y[j], alpha * acc);
    }
}

// ------------------------------------------------------------------------------------------
// fallback: thread per output, direct global loads
// ------------------------------------------------------------------------------------------
template <typename T, bool TRANS>
__global__ void __launch_bounds__(256) gbmv_direct_kernel(int64_t m, int64_t n, int64_t kl, int64_t ku, T alpha,
                                                         const T *__restrict__ A, int64_t lda,
                                                         const T *__restrict__ x, T beta, T *__restrict__ y) {
    const int64_t nout = TRANS ? n : m;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nout; o += stride) {
        T acc = 0;
        if (!TRANS) {
            int64_t jlo = o - kl > 0 ? o - kl : 0;
            int64_t jhi = o + ku < n - 1 ? o + ku : n - 1;
            for (int64_t j = jlo; j <= jhi; ++j) acc = fma(A[(ku + o - j) + lda * j], x[j], acc);
        } else {
            int64_t ilo = o - ku > 0 ? o - ku : 0;
            int64_t ihi = o + kl < m - 1 ? o + kl : m - 1;
            for (int64_t i = ilo; i <= ihi; ++i) acc = fma(A[(ku + i - o) + lda * o], x[i], acc);
        }
        y[o] = (beta == (T)0) ? alpha * acc : fma(beta, y[o], alpha * acc);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(int64_t n, T beta, T *__restrict__ y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += stride)
        y[o] = (beta == (T)0) ? (T)0 : beta * y[o];
}

// ------------------------------------------------------------------------------------------
// banded x dense (several right-hand sides):  Y[:, c] <- alpha*op(A)*X[:, c] + beta*Y[:, c]
// ------------------------------------------------------------------------------------------
// `mul!(C, A::BandedMatrix, B::Matrix, alpha, beta)`.  Same streamed-tile scheme as gbmv_tile_kernel, but a
// stage carries the slab of A ONCE plus the x-windows of R columns (R bulk copies), and a thread keeps R
// accumulators per output row -- so A is read from HBM once per R right-hand sides instead of once per column.
template <typename T>
struct GbmvMultiArgs {
    int64_t m, n, kl, ku, lda, nout;
    T alpha, beta;
    const T *A;
    const T *X;
    T *Y;
    int64_t ldx, ldy;
    int64_t tile, ntiles;
    int stages;
    int64_t sA_elems, sx_elems;   // per stage: sA_elems + R * sx_elems
};

template <typename T, int R, bool TRANS, int NBU>   // NBU: compile-time unroll bound >= kl+ku+1
__global__ void __launch_bounds__(GBMV_NT) gbmv_multi_kernel(const GbmvMultiArgs<T> a) {
    constexpr int EPV = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    T *sbase = reinterpret_cast<T *>(smem_raw + 128);
    const int tid = threadIdx.x;
    const int S = a.stages;
    const int nb = (int)(a.kl + a.ku + 1);
    const int64_t stage_elems = a.sA_elems + R * a.sx_elems;
    const int sxe = (int)a.sx_elems;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t first = blockIdx.x, step = gridDim.x;
    const int64_t mine = first < a.ntiles ? (a.ntiles - first + step - 1) / step : 0;
    const int64_t nin = TRANS ? a.m : a.n;

    auto ranges = [&](int64_t t, int64_t &o0, int64_t &o1, int64_t &c0, int64_t &c1, int64_t &x0, int64_t &x1) {
        o0 = t * a.tile;
        o1 = o0 + a.tile < a.nout ? o0 + a.tile : a.nout;
        const int64_t lo = TRANS ? a.ku : a.kl, hi = TRANS ? a.kl : a.ku;
        x0 = o0 - lo > 0 ? o0 - lo : 0;
        x1 = o1 + hi < nin ? o1 + hi : nin;
        if (x0 > nin) x0 = nin;
        if (x1 < x0) x1 = x0;
        if (!TRANS) { c0 = x0; c1 = x1; } else { c0 = o0; c1 = o1; }
    };
    auto issue = [&](int64_t t, int s) {  // thread 0 only
        int64_t o0, o1, c0, c1, x0, x1;
        ranges(t, o0, o1, c0, c1, x0, x1);
        const int64_t e0a = (c0 * a.lda) & ~(int64_t)(EPV - 1);
        const int64_t cntA = stage_round<T>(c1 * a.lda - e0a, a.n * a.lda - e0a);
        const int64_t x0a = x0 & ~(int64_t)(EPV - 1);
        const int64_t cntx = stage_round<T>(x1 - x0a, nin - x0a);
        T *sA = sbase + (int64_t)s * stage_elems;
        T *sx = sA + a.sA_elems;
        mbar_add_tx(&bars[s], stage_bytes_bulk<T>(cntA) + R * stage_bytes_bulk<T>(cntx));
        stage_issue<T>(sA, a.A + e0a, cntA, &bars[s]);
#pragma unroll
        for (int r = 0; r < R; ++r) stage_issue<T>(sx + r * a.sx_elems, a.X + r * a.ldx + x0a, cntx, &bars[s]);
        mbar_arrive(&bars[s]);
    };
    if (tid == 0) {
        const int64_t pro = mine < S ? mine : S;
        for (int64_t k = 0; k < pro; ++k) issue(first + k * step, (int)k);
    }
    const int lda = (int)a.lda;
    for (int64_t k = 0; k < mine; ++k) {
        const int s = (int)(k % S);
        const uint32_t parity = (uint32_t)((k / S) & 1);
        int64_t o0, o1, c0, c1, x0, x1;
        ranges(first + k * step, o0, o1, c0, c1, x0, x1);
        const int64_t e0a = (c0 * a.lda) & ~(int64_t)(EPV - 1);
        const int64_t x0a = x0 & ~(int64_t)(EPV - 1);
        const T *sA = sbase + (int64_t)s * stage_elems;
        const T *sx = sA + a.sA_elems;
        mbar_wait(&bars[s], parity);
        const bool interior = o0 - (TRANS ? a.ku : a.kl) >= 0 && o1 - 1 + (TRANS ? a.kl : a.ku) < nin;
        for (int64_t o = o0 + tid; o < o1; o += GBMV_NT) {
            T acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = (T)0;
            // N: terms j = o-kl+q, band row ku+o-j;  T: terms i = o-ku+q, band row ku+i-o.  Masked by index.
            const int64_t t0 = o - (TRANS ? a.ku : a.kl);
            const int xb = (int)(t0 - x0a);
            const int ab = TRANS ? (int)(o * a.lda - e0a) : (int)((a.kl + a.ku) + t0 * a.lda - e0a);
            const int astep = TRANS ? 1 : lda - 1;
            if (interior) {
#pragma unroll
                for (int q = 0; q < NBU; ++q) {
                    if (q < nb) {
                        const T av = sA[ab + q * astep];
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = fma(av, sx[r * sxe + xb + q], acc[r]);
                    }
                }
            } else {
                const int q0 = t0 < 0 ? (int)-t0 : 0;
                const int q1 = t0 + nb > nin ? (int)(nin - t0) : nb;
                for (int q = q0; q < q1; ++q) {
                    const T av = sA[ab + q * astep];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = fma(av, sx[r * sxe + xb + q], acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                T *yp = a.Y + r * a.ldy + o;
                *yp = (a.beta == (T)0) ? a.alpha * acc[r] : fma(a.beta, *yp, a.alpha * acc[r]);
            }
        }
        __syncthreads();
        if (tid == 0 && k + S < mine) issue(first + (k + S) * step, s);
    }
}

// ------------------------------------------------------------------------------------------
// fused broadcast sum:  y <- alpha1*A1*x + alpha2*A2*x + beta*y   (BroadcastArray(+/-, A1, A2) * x)
// ------------------------------------------------------------------------------------------
// `Mul{BroadcastBandedLayout{+}}` (reference src/LazyBandedMatrices.jl:20; used for D_xx +- D at
// test/test_misc.jl:21-40 and the Laplacian sum test/test_blockkron.jl:303,345) applied WITHOUT forming the sum
// and without a second pass over y: a stage carries both slabs and ONE x-window (the union band's).
template <typename T>
struct Gbmv2Args {
    int64_t m, n;
    int64_t kl[2], ku[2], lda[2];
    T alpha[2], beta;
    const T *A[2];
    const T *x;
    T *y;
    int64_t tile, ntiles;
    int stages;
    int64_t sA_elems[2], sx_elems;
};

template <typename T, int NBU>   // NBU: compile-time unroll bound >= both band widths
__global__ void __launch_bounds__(GBMV_NT) gbmv2_tile_kernel(const Gbmv2Args<T> a) {
    constexpr int EPV = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    T *sbase = reinterpret_cast<T *>(smem_raw + 128);
    const int tid = threadIdx.x;
    const int S = a.stages;
    const int64_t stage_elems = a.sA_elems[0] + a.sA_elems[1] + a.sx_elems;
    const int64_t klm = a.kl[0] > a.kl[1] ? a.kl[0] : a.kl[1], kum = a.ku[0] > a.ku[1] ? a.ku[0] : a.ku[1];
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t first = blockIdx.x, step = gridDim.x;
    const int64_t mine = first < a.ntiles ? (a.ntiles - first + step - 1) / step : 0;

    // column range [c0, c1) operand f needs for the output rows [o0, o1)
    auto cols = [&](int64_t o0, int64_t o1, int64_t kl, int64_t ku, int64_t &c0, int64_t &c1) {
        c0 = o0 - kl > 0 ? o0 - kl : 0;
        c1 = o1 + ku < a.n ? o1 + ku : a.n;
        if (c0 > a.n) c0 = a.n;
        if (c1 < c0) c1 = c0;
    };
    auto issue = [&](int64_t t, int s) {  // thread 0 only
        const int64_t o0 = t * a.tile, o1 = o0 + a.tile < a.m ? o0 + a.tile : a.m;
        T *dst = sbase + (int64_t)s * stage_elems;
        uint32_t bytes = 0;
        int64_t e0[2], cnt[2], x0, x1;
        for (int f = 0; f < 2; ++f) {
            int64_t c0, c1;
            cols(o0, o1, a.kl[f], a.ku[f], c0, c1);
            e0[f] = (c0 * a.lda[f]) & ~(int64_t)(EPV - 1);
            cnt[f] = c1 * a.lda[f] - e0[f];
            bytes += stage_bytes_bulk<T>(cnt[f]);
        }
        cols(o0, o1, klm, kum, x0, x1);
        const int64_t x0a = x0 & ~(int64_t)(EPV - 1);
        const int64_t cntx = x1 - x0a;
        bytes += stage_bytes_bulk<T>(cntx);
        mbar_add_tx(&bars[s], bytes);
        stage_issue<T>(dst, a.A[0] + e0[0], cnt[0], &bars[s]);
        stage_issue<T>(dst + a.sA_elems[0], a.A[1] + e0[1], cnt[1], &bars[s]);
        stage_issue<T>(dst + a.sA_elems[0] + a.sA_elems[1], a.x + x0a, cntx, &bars[s]);
        mbar_arrive(&bars[s]);
    };
    if (tid == 0) {
        const int64_t pro = mine < S ? mine : S;
        for (int64_t k = 0; k < pro; ++k) issue(first + k * step, (int)k);
    }
    for (int64_t k = 0; k < mine; ++k) {
        const int s = (int)(k % S);
        const uint32_t parity = (uint32_t)((k / S) & 1);
        const int64_t t = first + k * step;
        const int64_t o0 = t * a.tile, o1 = o0 + a.tile < a.m ? o0 + a.tile : a.m;
        const T *sA0 = sbase + (int64_t)s * stage_elems;
        const T *sA[2] = {sA0, sA0 + a.sA_elems[0]};
        const T *sx = sA0 + a.sA_elems[0] + a.sA_elems[1];
        int64_t e0[2], x0, x1;
        for (int f = 0; f < 2; ++f) {
            int64_t c0, c1;
            cols(o0, o1, a.kl[f], a.ku[f], c0, c1);
            e0[f] = (c0 * a.lda[f]) & ~(int64_t)(EPV - 1);
        }
        cols(o0, o1, klm, kum, x0, x1);
        const int64_t x0a = x0 & ~(int64_t)(EPV - 1);
        mbar_wait(&bars[s], parity);
        const bool interior = o0 - klm >= 0 && o1 - 1 + kum < a.n;   // no index masks needed
        for (int64_t o = o0 + tid; o < o1; o += GBMV_NT) {
            T tot = (T)0;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const int nb = (int)(a.kl[f] + a.ku[f] + 1), lda = (int)a.lda[f];
                const int64_t j0 = o - a.kl[f];
                const int xb = (int)(j0 - x0a);
                const int ab = (int)((a.kl[f] + a.ku[f]) + j0 * a.lda[f] - e0[f]);
                T acc = (T)0;
                if (interior) {
#pragma unroll
                    for (int q = 0; q < NBU; ++q)
                        if (q < nb) acc = fma(sA[f][ab + q * (lda - 1)], sx[xb + q], acc);
                } else {
                    const int q0 = j0 < 0 ? (int)-j0 : 0;
                    const int q1 = j0 + nb > a.n ? (int)(a.n - j0) : nb;
                    for (int q = q0; q < q1; ++q) acc = fma(sA[f][ab + q * (lda - 1)], sx[xb + q], acc);
                }
                tot = fma(a.alpha[f], acc, tot);
            }
            a.y[o] = (a.beta == (T)0) ? tot : fma(a.beta, a.y[o], tot);
        }
        __syncthreads();
        if (tid == 0 && k + S < mine) issue(first + (k + S) * step, s);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }

template <typename T, int NB, bool TRANS, bool HALO>
int launch_tile(lbm_handle_t h, const GbmvArgs<T> &a, int grid, size_t smem, const HaloFuse &hf) {
    static std::atomic<int> configured_dev_mask{0};
    auto kern = gbmv_tile_kernel<T, NB, TRANS, HALO>;
    if (!(configured_dev_mask.load(std::memory_order_acquire) & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
        configured_dev_mask.fetch_or(1 << h->device, std::memory_order_release);
    }
    LBM_CUDA(h, lbm_launch_pdl(h, kern, dim3(grid), dim3(GBMV_NT), smem, a, hf));
    LBM_LAUNCH_CHECK(h, "gbmv_tile_kernel");
    return 0;
}

template <typename T, bool TRANS, bool HALO>
int dispatch_tile(lbm_handle_t h, const GbmvArgs<T> &a, int grid, size_t smem, const HaloFuse &hf) {
    switch ((int)(a.kl + a.ku + 1)) {
        case 1: return launch_tile<T, 1, TRANS, HALO>(h, a, grid, smem, hf);
        case 2: return launch_tile<T, 2, TRANS, HALO>(h, a, grid, smem, hf);
        case 3: return launch_tile<T, 3, TRANS, HALO>(h, a, grid, smem, hf);
        case 5: return launch_tile<T, 5, TRANS, HALO>(h, a, grid, smem, hf);
        case 7: return launch_tile<T, 7, TRANS, HALO>(h, a, grid, smem, hf);
        case 9: return launch_tile<T, 9, TRANS, HALO>(h, a, grid, smem, hf);
        case 17: return launch_tile<T, 17, TRANS, HALO>(h, a, grid, smem, hf);
        case 33: return launch_tile<T, 33, TRANS, HALO>(h, a, grid, smem, hf);
        default: return launch_tile<T, 0, TRANS, HALO>(h, a, grid, smem, hf);
    }
}

// `hf` != NULL: the caller (dist.cu) wants the ghost exchange fused into this launch; only the streamed-tile kernel can
// do that, so any other route returns LBM_NOT_FUSED WITHOUT launching and the caller takes the unfused path.
template <typename T>
int gbmv_impl(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, T alpha, const T *A,
              int64_t lda, const T *x, T beta, T *y, const HaloFuse *hf = nullptr) {
    if (!h) return -1;
    const bool tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!tr && !(trans == 'N' || trans == 'n')) return lbm_fail_arg(h, 2, "trans must be N, T or C");
    if (m < 0) return lbm_fail_arg(h, 3, "m < 0");
    if (n < 0) return lbm_fail_arg(h, 4, "n < 0");
    if (kl < 0) return lbm_fail_arg(h, 5, "kl < 0");
    if (ku < 0) return lbm_fail_arg(h, 6, "ku < 0");
    if (lda < kl + ku + 1) return lbm_fail_arg(h, 9, "lda < kl+ku+1");
    const int64_t nout = tr ? n : m;
    const int64_t nin = tr ? m : n;
    if (nout == 0) return 0;
    if (!y) return lbm_fail_arg(h, 12, "y == NULL");
    lbm_device_guard g(h->device);
    constexpr int EPV = 16 / (int)sizeof(T);

    if (hf && (nin == 0 || alpha == (T)0 || tr)) return LBM_NOT_FUSED;
    if (nin == 0 || alpha == (T)0) {  // y <- beta*y
        int64_t blocks = (nout + 255) / 256;
        if (blocks > (int64_t)h->sm_count * 8) blocks = (int64_t)h->sm_count * 8;
        scale_kernel<T><<<(unsigned)blocks, 256, 0, h->stream>>>(nout, beta, y);
        LBM_LAUNCH_CHECK(h, "scale_kernel");
        return 0;
    }
    if (!A) return lbm_fail_arg(h, 8, "A == NULL");
    if (!x) return lbm_fail_arg(h, 10, "x == NULL");

    const int64_t nb = kl + ku + 1;
    const bool aligned = lbm_aligned16(A) && lbm_aligned16(x);
    const int64_t smem_cap = h->max_smem_optin - 1024;
    const int64_t force = h->tune.gbmv_force;

    // ---- narrow streamed-tile path ------------------------------------------------------
    // feasible when a 256-output tile (+halo columns) of the slab fits one ~100 KB stage
    const int64_t min_stage_bytes = ((GBMV_NT + nb) * (lda + 1) + 2 * EPV) * (int64_t)sizeof(T);
    bool narrow_ok = aligned && min_stage_bytes <= 100 * 1024 && lda <= 4 * nb + 8;
    bool wide_ok = aligned && !tr && (nb + 32) <= 1024 && (32 * lda + 64) * (int64_t)sizeof(T) <= smem_cap;
    int path;  // 1 direct, 2 narrow, 3 wide-n, 4 wide-t (warp per output, global), 5 wide-t (TMA ring)
    if (force == 1) path = 1;
    else if (force == 2 && narrow_ok) path = 2;
    else if (force == 3 && wide_ok) path = 3;
    else if (narrow_ok) path = 2;
    else if (wide_ok) path = 3;
    else if (tr && nb >= 32 && aligned && force != 4 && (32 * lda + 64 + 32 + nb) * (int64_t)sizeof(T) <= smem_cap) path = 5;
    else if (tr && nb >= 32) path = 4;
    else path = 1;
    if (hf && path != 2) return LBM_NOT_FUSED;

    if (path == 2) {
        GbmvArgs<T> a;
        a.m = m; a.n = n; a.kl = kl; a.ku = ku; a.lda = lda; a.nout = nout;
        a.alpha = alpha; a.beta = beta; a.A = A; a.x = x; a.y = y;
        // Tuned on B200 (profiles/sweep_r01.jsonl): ~33-39 KB stages, 2-deep ring, and as many
        // CTAs per SM as keep ~200 KB of bulk copies in flight per SM.
        a.skew = h->tune.gbmv_skew > 0 ? round_up(h->tune.gbmv_skew, EPV) : h->tune.gbmv_skew;
        const int64_t slack = a.skew > 0 ? a.skew : 0;
        auto stage_elems = [&](int64_t tl, int64_t &sa, int64_t &sx) {
            sa = round_up((tl + kl + ku) * lda + EPV + slack, EPV);
            sx = round_up(tl + kl + ku + EPV + slack, EPV);
        };
        int64_t sa, sx;
        int64_t tile = h->tune.gbmv_tile;
        // Three stored diagonals, Float64 (the metric's l = u = 1 point and every rank's (0,2) / (2,0) slab of its row shards):
        // FOUR single-stage CTAs per SM with 1280-row tiles instead of three 2-stage ones.  Measured on one B200 at n = 2^24..2^27
        // (benchmarks/rect_probe.py, profiles/rect_probe_r02.md): the default was 3-8 % slower on the (0,2) / (2,0) slabs -- whose
        // bulk copies start on exact 8 KB multiples -- than on (1,1); this setting is within 1 % of the best for all three shapes.
        const bool three_diag = sizeof(T) == 8 && nb <= 3 && h->tune.gbmv_tile <= 0 && h->tune.gbmv_stages <= 0 && h->tune.gbmv_ctas_per_sm <= 0;
        if (three_diag) {
            tile = 1280;
            // one rank of an 8-GPU job holds 2^24 rows: 13108 tiles of 1280 over 592 CTAs is 22.1 per CTA, i.e. 23 for some
            // (4 % of imbalance); 1024-row tiles give 27.7 -> 28 (1 %).  Measured on the (0,2) / (2,0) slabs at that size:
            // 103.5 -> 100.7 us; the (1,1) slab of rank 0 prefers 1280 at every size (rect_probe_r02.md), so only rectangular
            // slabs switch, and only when the tile count favours it by more than 1.5 %.
            if (kl != ku) {
                const int64_t ctas4 = (int64_t)h->sm_count * 4;
                auto cost = [&](int64_t tl) { const int64_t nt = (nout + tl - 1) / tl; return ((nt + ctas4 - 1) / ctas4) * tl; };
                if (cost(1024) * 1000 < cost(1280) * 985) tile = 1024;
            }
        }
        if (tile <= 0) {
            tile = GBMV_NT;
            for (int64_t tl = 2 * GBMV_NT; tl <= 2048; tl += GBMV_NT) {
                stage_elems(tl, sa, sx);
                if ((sa + sx) * (int64_t)sizeof(T) <= 36 * 1024) tile = tl;
            }
        }
        tile = round_up(tile, GBMV_NT);
        stage_elems(tile, sa, sx);
        while (tile > GBMV_NT && (sa + sx) * (int64_t)sizeof(T) + 128 > smem_cap) {
            tile -= GBMV_NT;
            stage_elems(tile, sa, sx);
        }
        a.tile = tile;
        a.ntiles = (nout + tile - 1) / tile;
        a.sA_elems = sa;
        a.sx_elems = sx;
        const int64_t stage_bytes = (sa + sx) * (int64_t)sizeof(T);
        int64_t stages = h->tune.gbmv_stages > 0 ? h->tune.gbmv_stages : (three_diag ? 1 : 2);
        if (stages > GBMV_MAX_STAGES) stages = GBMV_MAX_STAGES;
        int64_t ctas = h->tune.gbmv_ctas_per_sm > 0 ? h->tune.gbmv_ctas_per_sm
                                                    : (three_diag ? 4 : (200 * 1024) / (stages * stage_bytes + 256));
        if (ctas < 1) ctas = 1;
        if (ctas > 8) ctas = 8;
        // fit ctas * (128 + stages*stage_bytes) into the SM's shared memory
        const int64_t sm_smem = 227 * 1024;
        while (stages > 1 && ctas * (stages * stage_bytes + 128 + 1024) > sm_smem) --stages;
        while (ctas > 1 && ctas * (stages * stage_bytes + 128 + 1024) > sm_smem) --ctas;
        a.stages = (int)stages;
        int64_t grid = (int64_t)h->sm_count * ctas;
        if (grid > a.ntiles) grid = a.ntiles;
        const size_t smem = (size_t)(128 + stages * stage_bytes);
        HaloFuse none;
        none.on = 0;
        if (hf) return dispatch_tile<T, false, true>(h, a, (int)grid, smem, *hf);
        return tr ? dispatch_tile<T, true, false>(h, a, (int)grid, smem, none) : dispatch_tile<T, false, false>(h, a, (int)grid, smem, none);
    }

    if (path == 3) {
        GbmvWideArgs<T> a;
        a.m = m; a.n = n; a.kl = kl; a.ku = ku; a.lda = lda;
        a.alpha = alpha; a.beta = beta; a.A = A; a.x = x; a.y = y;
        int64_t chunk = h->tune.gbmvw_chunk;
        if (chunk <= 0) {
            chunk = (64 * 1024) / (lda * (int64_t)sizeof(T));
            chunk = chunk / 32 * 32;
            if (chunk < 32) chunk = 32;
        }
        chunk = round_up(chunk, 32);
        while (chunk > 32 && (chunk + kl + ku > 1024 || (chunk * (lda + 1) + 2 * EPV) * (int64_t)sizeof(T) + 128 > smem_cap))
            chunk -= 32;
        const int64_t nt = round_up(chunk + kl + ku, 32);
        a.chunk = (int)chunk;
        a.sA_elems = round_up(chunk * lda + EPV, EPV);
        a.sx_elems = round_up(chunk + EPV, EPV);
        const int64_t stage_bytes = (a.sA_elems + a.sx_elems) * (int64_t)sizeof(T);
        int64_t stages = h->tune.gbmvw_stages > 0 ? h->tune.gbmvw_stages : 3;  // 2..6 within noise on B200
        if (stages > GBMV_MAX_STAGES) stages = GBMV_MAX_STAGES;
        while (stages > 1 && stages * stage_bytes + 128 > smem_cap) --stages;
        a.stages = (int)stages;
        const int64_t smem_cta = stages * stage_bytes + 128;
        int64_t ctas = h->tune.gbmvw_ctas_per_sm > 0 ? h->tune.gbmvw_ctas_per_sm : (227 * 1024) / (smem_cta + 1024);
        if (ctas < 1) ctas = 1;
        if (ctas * nt > 2048) ctas = 2048 / nt;
        if (ctas < 1) ctas = 1;
        int64_t grid = (int64_t)h->sm_count * ctas;
        int64_t run = (m + grid - 1) / grid;
        const int64_t min_run = 4 * (kl + ku) > 1024 ? 4 * (kl + ku) : 1024;  // keep halo re-reads <= 25 %
        if (run < min_run) run = min_run;
        run = round_up(run, EPV);
        grid = (m + run - 1) / run;
        a.run = run;
        static std::atomic<int> configured_dev_mask{0};
        auto kern = gbmv_wide_n_kernel<T>;
        if (!(configured_dev_mask & (1 << h->device))) {
            LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
            configured_dev_mask |= (1 << h->device);
        }
        kern<<<(unsigned)grid, (unsigned)nt, (size_t)smem_cta, h->stream>>>(a);
        LBM_LAUNCH_CHECK(h, "gbmv_wide_n_kernel");
        return 0;
    }

    if (path == 5) {
        GbmvWideArgs<T> a;
        a.m = m; a.n = n; a.kl = kl; a.ku = ku; a.lda = lda;
        a.alpha = alpha; a.beta = beta; a.A = A; a.x = x; a.y = y;
        int64_t chunk = h->tune.gbmvw_chunk > 0 ? h->tune.gbmvw_chunk : 32;
        chunk = round_up(chunk, 32);
        if (chunk > 256) chunk = 256;
        auto sbytes = [&](int64_t c) { return (round_up(c * lda + EPV, EPV) + round_up(c + nb + 2 * EPV, EPV)) * (int64_t)sizeof(T); };
        while (chunk > 32 && sbytes(chunk) + 128 > smem_cap) chunk -= 32;
        a.chunk = (int)chunk;
        a.sA_elems = round_up(chunk * lda + EPV, EPV);
        a.sx_elems = round_up(chunk + nb + 2 * EPV, EPV);
        const int64_t stage_bytes = sbytes(chunk);
        // one CTA per SM with 66 KB chunks: the ring must keep >= 2 chunks in flight behind the one
        // being reduced, i.e. 3 stages
        int64_t stages = h->tune.gbmvw_stages > 0 ? h->tune.gbmvw_stages : 3;
        if (stages > GBMV_MAX_STAGES) stages = GBMV_MAX_STAGES;
        while (stages > 1 && stages * stage_bytes + 128 > smem_cap) --stages;
        a.stages = (int)stages;
        while (stages > 1 && stages * stage_bytes + 128 + 8 * chunk * (int64_t)sizeof(T) > smem_cap) --stages;
        a.stages = (int)stages;
        const int64_t smem_cta = stages * stage_bytes + 128 + 8 * chunk * (int64_t)sizeof(T);
        int64_t ctas = h->tune.gbmvw_ctas_per_sm > 0 ? h->tune.gbmvw_ctas_per_sm : (227 * 1024) / (smem_cta + 1024);
        if (ctas < 1) ctas = 1;
        if (ctas > 8) ctas = 8;
        int64_t grid = (int64_t)h->sm_count * ctas;
        int64_t run = (n + grid - 1) / grid;
        if (run < chunk) run = chunk;
        run = round_up(run, 32);
        grid = (n + run - 1) / run;
        a.run = run;
        static std::atomic<int> configured_dev_mask{0};
        auto kern = gbmv_wide_t_ring_kernel<T>;
        if (!(configured_dev_mask & (1 << h->device))) {
            LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
            configured_dev_mask |= (1 << h->device);
        }
        kern<<<(unsigned)grid, 256, (size_t)smem_cta, h->stream>>>(a);
        LBM_LAUNCH_CHECK(h, "gbmv_wide_t_ring_kernel");
        return 0;
    }

    if (path == 4) {
        int64_t blocks = (n + 7) / 8;  // 8 warps per block, one output per warp
        if (blocks > (int64_t)h->sm_count * 8) blocks = (int64_t)h->sm_count * 8;
        gbmv_wide_t_kernel<T><<<(unsigned)blocks, 256, 0, h->stream>>>(m, n, kl, ku, alpha, A, lda, x, beta, y);
        LBM_LAUNCH_CHECK(h, "gbmv_wide_t_kernel");
        return 0;
    }

    {
        int64_t blocks = (nout + 255) / 256;
        if (blocks > (int64_t)h->sm_count * 8) blocks = (int64_t)h->sm_count * 8;
        if (tr)
            gbmv_direct_kernel<T, true><<<(unsigned)blocks, 256, 0, h->stream>>>(m, n, kl, ku, alpha, A, lda, x, beta, y);
        else
            gbmv_direct_kernel<T, false><<<(unsigned)blocks, 256, 0, h->stream>>>(m, n, kl, ku, alpha, A, lda, x, beta, y);
        LBM_LAUNCH_CHECK(h, "gbmv_direct_kernel");
        return 0;
    }
}

template <typename T, int R, bool TRANS, int NBU>
int launch_multi2(lbm_handle_t h, const GbmvMultiArgs<T> &a, int grid, size_t smem) {
    static std::atomic<int> configured_dev_mask{0};
    auto kern = gbmv_multi_kernel<T, R, TRANS, NBU>;
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
        configured_dev_mask |= (1 << h->device);
    }
    kern<<<grid, GBMV_NT, smem, h->stream>>>(a);
    LBM_LAUNCH_CHECK(h, "gbmv_multi_kernel");
    return 0;
}

template <typename T, int R, bool TRANS>
int launch_multi(lbm_handle_t h, const GbmvMultiArgs<T> &a, int grid, size_t smem) {
    const int64_t nb = a.kl + a.ku + 1;
    if (nb <= 3) return launch_multi2<T, R, TRANS, 3>(h, a, grid, smem);
    if (nb <= 17) return launch_multi2<T, R, TRANS, 17>(h, a, grid, smem);
    return launch_multi2<T, R, TRANS, 33>(h, a, grid, smem);
}

// Y <- alpha*op(A)*X + beta*Y, X/Y column-major with leading dimensions ldx/ldy.  Narrow bands: groups of up to 8
// columns share one pass over A (gbmv_multi_kernel); wide bands / unaligned operands: one gbmv per column.
template <typename T>
int gbmv_multi_impl(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, T alpha, const T *A,
                    int64_t lda, const T *X, int64_t ldx, T beta, T *Y, int64_t ldy, int64_t nrhs) {
    if (!h) return -1;
    const bool tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!tr && !(trans == 'N' || trans == 'n')) return lbm_fail_arg(h, 2, "trans must be N, T or C");
    if (m < 0) return lbm_fail_arg(h, 3, "m < 0");
    if (n < 0) return lbm_fail_arg(h, 4, "n < 0");
    if (kl < 0) return lbm_fail_arg(h, 5, "kl < 0");
    if (ku < 0) return lbm_fail_arg(h, 6, "ku < 0");
    if (lda < kl + ku + 1) return lbm_fail_arg(h, 9, "lda < kl+ku+1");
    if (nrhs < 0) return lbm_fail_arg(h, 15, "nrhs < 0");
    const int64_t nout = tr ? n : m, nin = tr ? m : n;
    if (nrhs > 1 && ldx < nin) return lbm_fail_arg(h, 11, "ldx < rows of X");
    if (nrhs > 1 && ldy < nout) return lbm_fail_arg(h, 14, "ldy < rows of Y");
    if (nrhs == 0 || nout == 0) return 0;
    constexpr int EPV = 16 / (int)sizeof(T);
    const int64_t nb = kl + ku + 1;
    const int64_t smem_cap = h->max_smem_optin - 1024;
    const bool fused_ok = nrhs >= 2 && nin > 0 && alpha != (T)0 && A && X && Y && nb <= 33 && lda <= 4 * nb + 8 &&
                          lbm_aligned16(A) && lbm_aligned16(X) && lbm_aligned16(Y) && ldx % EPV == 0 &&
                          h->tune.gbmv_force != 1;
    int64_t c = 0;
    if (fused_ok) {
        lbm_device_guard g(h->device);
        while (nrhs - c >= 2) {
            const int R = nrhs - c >= 8 ? 8 : (nrhs - c >= 4 ? 4 : 2);
            GbmvMultiArgs<T> a;
            a.m = m; a.n = n; a.kl = kl; a.ku = ku; a.lda = lda; a.nout = nout; a.alpha = alpha; a.beta = beta;
            a.A = A; a.X = X + c * ldx; a.Y = Y + c * ldy; a.ldx = ldx; a.ldy = ldy;
            auto stage_elems = [&](int64_t tl, int64_t &sa, int64_t &sx) {
                sa = round_up((tl + kl + ku) * lda + 2 * EPV, EPV);   // + alignment offset + stage_round slack
                sx = round_up(tl + kl + ku + 2 * EPV, EPV);
            };
            int64_t sa, sx, tile = GBMV_NT;
            for (int64_t tl = 2 * GBMV_NT; tl <= 2048; tl += GBMV_NT) {
                stage_elems(tl, sa, sx);
                if ((sa + R * sx) * (int64_t)sizeof(T) <= 40 * 1024) tile = tl;
            }
            if (h->tune.gbmv_tile > 0) tile = round_up(h->tune.gbmv_tile, GBMV_NT);
            stage_elems(tile, sa, sx);
            while (tile > GBMV_NT && 2 * (sa + R * sx) * (int64_t)sizeof(T) + 128 > smem_cap) {
                tile -= GBMV_NT;
                stage_elems(tile, sa, sx);
            }
            a.tile = tile; a.ntiles = (nout + tile - 1) / tile; a.sA_elems = sa; a.sx_elems = sx;
            const int64_t stage_bytes = (sa + R * sx) * (int64_t)sizeof(T);
            if (stage_bytes + 128 > smem_cap) break;       // cannot happen for nb <= 33; falls back below
            int64_t stages = 2;
            int64_t ctas = (227 * 1024) / (stages * stage_bytes + 128 + 1024);   // as many CTAs as the SM's smem holds
            if (ctas < 1) ctas = 1;
            if (ctas > 8) ctas = 8;
            while (stages > 1 && ctas * (stages * stage_bytes + 128 + 1024) > 227 * 1024) --stages;
            while (ctas > 1 && ctas * (stages * stage_bytes + 128 + 1024) > 227 * 1024) --ctas;
            a.stages = (int)stages;
            int64_t grid = (int64_t)h->sm_count * ctas;
            if (grid > a.ntiles) grid = a.ntiles;
            const size_t smem = (size_t)(128 + stages * stage_bytes);
            int rc;
            if (R == 8) rc = tr ? launch_multi<T, 8, true>(h, a, (int)grid, smem) : launch_multi<T, 8, false>(h, a, (int)grid, smem);
            else if (R == 4) rc = tr ? launch_multi<T, 4, true>(h, a, (int)grid, smem) : launch_multi<T, 4, false>(h, a, (int)grid, smem);
            else rc = tr ? launch_multi<T, 2, true>(h, a, (int)grid, smem) : launch_multi<T, 2, false>(h, a, (int)grid, smem);
            if (rc) return rc;
            c += R;
        }
    }
    // wide bands, op = N: groups of 8 / 4 / 2 columns share one pass over A (gbmv_wide_n_multi_kernel)
    {
        const bool wide_multi_ok = !tr && nb > 33 && nin > 0 && alpha != (T)0 && A && X && Y && lbm_aligned16(A) && lbm_aligned16(X) &&
                                   lbm_aligned16(Y) && ldx % EPV == 0 && (nb + 32) <= 1024 && h->tune.gbmv_force != 1 &&
                                   (32 * (lda + 8) + 64) * (int64_t)sizeof(T) <= smem_cap;
        if (wide_multi_ok) {
            lbm_device_guard g(h->device);
            while (nrhs - c >= 2) {
                const int R = nrhs - c >= 8 ? 8 : (nrhs - c >= 4 ? 4 : 2);
                GbmvWideMultiArgs<T> am;
                GbmvWideArgs<T> &a = am.w;
                a.m = m; a.n = n; a.kl = kl; a.ku = ku; a.lda = lda;
                a.alpha = alpha; a.beta = beta; a.A = A; a.x = X + c * ldx; a.y = Y + c * ldy;
                am.ldx = ldx; am.ldy = ldy;
                int64_t chunk = h->tune.gbmvw_chunk > 0 ? h->tune.gbmvw_chunk : (64 * 1024) / (lda * (int64_t)sizeof(T));
                chunk = chunk / 16 * 16;
                if (chunk < 16) chunk = 16;
                while (chunk > 16 && (chunk + kl + ku > 1024 || (chunk * (lda + R) + (R + 1) * 2 * EPV) * (int64_t)sizeof(T) + 128 > smem_cap))
                    chunk -= 16;
                const int64_t nt = round_up(chunk + kl + ku, 32);
                a.chunk = (int)chunk;
                a.sA_elems = round_up(chunk * lda + 2 * EPV, EPV);
                a.sx_elems = round_up(chunk + 2 * EPV, EPV);
                const int64_t stage_bytes = (a.sA_elems + R * a.sx_elems) * (int64_t)sizeof(T);
                int64_t stages = h->tune.gbmvw_stages > 0 ? h->tune.gbmvw_stages : 3;
                if (stages > GBMV_MAX_STAGES) stages = GBMV_MAX_STAGES;
                while (stages > 1 && stages * stage_bytes + 128 > smem_cap) --stages;
                a.stages = (int)stages;
                const int64_t smem_cta = stages * stage_bytes + 128;
                int64_t ctas = (227 * 1024) / (smem_cta + 1024);
                if (ctas < 1) ctas = 1;
                if (ctas * nt > 2048) ctas = 2048 / nt;
                if (ctas < 1) ctas = 1;
                int64_t grid = (int64_t)h->sm_count * ctas;
                int64_t run = (m + grid - 1) / grid;
                const int64_t min_run = 4 * (kl + ku) > 1024 ? 4 * (kl + ku) : 1024;
                if (run < min_run) run = min_run;
                run = round_up(run, EPV);
                grid = (m + run - 1) / run;
                a.run = run;
#define LBM_WIDE_MULTI(RR)                                                                                              \
    do {                                                                                                                \
        static std::atomic<int> cfg2[2] = {};                                                                                  \
        std::atomic<int> &cfg = cfg2[nt <= 512 ? 0 : 1];                                                                          \
        auto kern = nt <= 512 ? gbmv_wide_n_multi_kernel<T, RR, 512> : gbmv_wide_n_multi_kernel<T, RR, 1024>;           \
        if (!(cfg & (1 << h->device))) {                                                                                \
            LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));    \
            cfg |= (1 << h->device);                                                                                    \
        }                                                                                                               \
        kern<<<(unsigned)grid, (unsigned)nt, (size_t)smem_cta, h->stream>>>(am);                                        \
    } while (0)
                if (R == 8) LBM_WIDE_MULTI(8);
                else if (R == 4) LBM_WIDE_MULTI(4);
                else LBM_WIDE_MULTI(2);
#undef LBM_WIDE_MULTI
                LBM_LAUNCH_CHECK(h, "gbmv_wide_n_multi_kernel");
                c += R;
            }
        }
    }
    // wide bands, op = T: groups of 8 / 4 / 2 columns share one pass over A (gbmv_wide_t_multi_kernel)
    {
        const bool wide_t_multi_ok = tr && nb > 33 && nin > 0 && n > 0 && alpha != (T)0 && A && X && Y && lbm_aligned16(A) && lbm_aligned16(X) &&
                                     ldx % EPV == 0 && h->tune.gbmv_force != 1 && h->tune.gbmv_force != 4;
        if (wide_t_multi_ok) {
            lbm_device_guard g(h->device);
            while (nrhs - c >= 2) {
                const int R = nrhs - c >= 8 ? 8 : (nrhs - c >= 4 ? 4 : 2);
                GbmvWideMultiArgs<T> am;
                GbmvWideArgs<T> &a = am.w;
                a.m = m; a.n = n; a.kl = kl; a.ku = ku; a.lda = lda;
                a.alpha = alpha; a.beta = beta; a.A = A; a.x = X + c * ldx; a.y = Y + c * ldy;
                am.ldx = ldx; am.ldy = ldy;
                int64_t chunk = h->tune.gbmvw_chunk > 0 ? round_up(h->tune.gbmvw_chunk, 32) : 32;
                if (chunk > 256) chunk = 256;
                auto sA_el = [&](int64_t ch) { return round_up(ch * lda + 2 * EPV, EPV); };
                auto sx_el = [&](int64_t ch) { return round_up(ch + nb + 3 * EPV, EPV); };
                auto sbytes = [&](int64_t ch) { return (sA_el(ch) + R * sx_el(ch)) * (int64_t)sizeof(T); };
                while (chunk > 32 && sbytes(chunk) + 128 + 8 * R * chunk * (int64_t)sizeof(T) > smem_cap) chunk -= 32;
                a.chunk = (int)chunk;
                a.sA_elems = sA_el(chunk);
                a.sx_elems = sx_el(chunk);
                const int64_t stage_bytes = sbytes(chunk);
                const int64_t part_bytes = 8 * R * chunk * (int64_t)sizeof(T);
                int64_t stages = h->tune.gbmvw_stages > 0 ? h->tune.gbmvw_stages : 3;
                if (stages > GBMV_MAX_STAGES) stages = GBMV_MAX_STAGES;
                while (stages > 1 && stages * stage_bytes + 128 + part_bytes > smem_cap) --stages;
                a.stages = (int)stages;
                const int64_t smem_cta = stages * stage_bytes + 128 + part_bytes;
                if (smem_cta > smem_cap) break;   // band too wide for R windows + a 32-column slab: column by column below
                int64_t ctas = h->tune.gbmvw_ctas_per_sm > 0 ? h->tune.gbmvw_ctas_per_sm : (227 * 1024) / (smem_cta + 1024);
                if (ctas < 1) ctas = 1;
                if (ctas > 8) ctas = 8;
                int64_t grid = (int64_t)h->sm_count * ctas;
                int64_t run = (n + grid - 1) / grid;
                if (run < chunk) run = chunk;
                run = round_up(run, 32);
                grid = (n + run - 1) / run;
                a.run = run;
#define LBM_WIDE_T_MULTI(RR)                                                                                            \
    do {                                                                                                                \
        static std::atomic<int> cfg{0};                                                                                 \
        auto kern = gbmv_wide_t_multi_kernel<T, RR>;                                                                    \
        if (!(cfg & (1 << h->device))) {                                                                                \
            LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));    \
            cfg |= (1 << h->device);                                                                                    \
        }                                                                                                               \
        kern<<<(unsigned)grid, 256, (size_t)smem_cta, h->stream>>>(am);                                                 \
    } while (0)
                if (R == 8) LBM_WIDE_T_MULTI(8);
                else if (R == 4) LBM_WIDE_T_MULTI(4);
                else LBM_WIDE_T_MULTI(2);
#undef LBM_WIDE_T_MULTI
                LBM_LAUNCH_CHECK(h, "gbmv_wide_t_multi_kernel");
                c += R;
            }
        }
    }
    for (; c < nrhs; ++c) {
        int rc = gbmv_impl<T>(h, trans, m, n, kl, ku, alpha, A, lda, X ? X + c * ldx : X, beta, Y ? Y + c * ldy : Y);
        if (rc) return rc;
    }
    return 0;
}

template <typename T, int NBU>
int launch_gbmv2(lbm_handle_t h, const Gbmv2Args<T> &a, int grid, size_t smem) {
    static std::atomic<int> configured_dev_mask{0};
    auto kern = gbmv2_tile_kernel<T, NBU>;
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
        configured_dev_mask |= (1 << h->device);
    }
    kern<<<grid, GBMV_NT, smem, h->stream>>>(a);
    LBM_LAUNCH_CHECK(h, "gbmv2_tile_kernel");
    return 0;
}

// y <- alpha1*A1*x + alpha2*A2*x + beta*y.  Fused when both bands are narrow and everything is 16-byte aligned;
// otherwise two gbmv calls (the second accumulating with beta = 1).
template <typename T>
int gbmv2_impl(lbm_handle_t h, int64_t m, int64_t n, T alpha1, int64_t kl1, int64_t ku1, const T *A1, int64_t lda1, T alpha2,
               int64_t kl2, int64_t ku2, const T *A2, int64_t lda2, const T *x, T beta, T *y) {
    if (!h) return -1;
    if (m < 0) return lbm_fail_arg(h, 2, "m < 0");
    if (n < 0) return lbm_fail_arg(h, 3, "n < 0");
    if (kl1 < 0 || ku1 < 0) return lbm_fail_arg(h, 5, "kl1/ku1 < 0");
    if (lda1 < kl1 + ku1 + 1) return lbm_fail_arg(h, 8, "lda1 < kl1+ku1+1");
    if (kl2 < 0 || ku2 < 0) return lbm_fail_arg(h, 10, "kl2/ku2 < 0");
    if (lda2 < kl2 + ku2 + 1) return lbm_fail_arg(h, 13, "lda2 < kl2+ku2+1");
    if (m == 0) return 0;
    constexpr int EPV = 16 / (int)sizeof(T);
    const int64_t nb1 = kl1 + ku1 + 1, nb2 = kl2 + ku2 + 1;
    const int64_t klm = kl1 > kl2 ? kl1 : kl2, kum = ku1 > ku2 ? ku1 : ku2;
    auto stage_elems = [&](int64_t tl, int64_t &s1, int64_t &s2, int64_t &sx) {
        s1 = round_up((tl + kl1 + ku1) * lda1 + EPV, EPV);
        s2 = round_up((tl + kl2 + ku2) * lda2 + EPV, EPV);
        sx = round_up(tl + klm + kum + EPV, EPV);
    };
    int64_t s1, s2, sx;
    stage_elems(GBMV_NT, s1, s2, sx);
    const bool fused = n > 0 && A1 && A2 && x && y && alpha1 != (T)0 && alpha2 != (T)0 && nb1 <= 33 && nb2 <= 33 &&
                       lda1 <= 4 * nb1 + 8 && lda2 <= 4 * nb2 + 8 && lbm_aligned16(A1) && lbm_aligned16(A2) && lbm_aligned16(x) &&
                       (s1 + s2 + sx) * (int64_t)sizeof(T) <= 100 * 1024 && h->tune.gbmv_force != 1;
    if (!fused) {
        int rc = gbmv_impl<T>(h, 'N', m, n, kl1, ku1, alpha1, A1, lda1, x, beta, y);
        if (rc) return rc;
        rc = gbmv_impl<T>(h, 'N', m, n, kl2, ku2, alpha2, A2, lda2, x, (T)1, y);
        return rc < 0 ? (rc <= -5 ? rc - 5 : rc) : rc;   // argument numbers of the second operand
    }
    lbm_device_guard g(h->device);
    Gbmv2Args<T> a;
    a.m = m; a.n = n; a.kl[0] = kl1; a.ku[0] = ku1; a.lda[0] = lda1; a.kl[1] = kl2; a.ku[1] = ku2; a.lda[1] = lda2;
    a.alpha[0] = alpha1; a.alpha[1] = alpha2; a.beta = beta; a.A[0] = A1; a.A[1] = A2; a.x = x; a.y = y;
    int64_t tile = GBMV_NT;
    for (int64_t tl = 2 * GBMV_NT; tl <= 2048; tl += GBMV_NT) {
        stage_elems(tl, s1, s2, sx);
        // two CTAs per SM with the largest 2-stage ring that allows: swept in profiles/configs_r01_final.jsonl ("BC fused tile=")
        if ((s1 + s2 + sx) * (int64_t)sizeof(T) <= 56 * 1024 + 512) tile = tl;
    }
    if (h->tune.gbmv_tile > 0) tile = round_up(h->tune.gbmv_tile, GBMV_NT);
    stage_elems(tile, s1, s2, sx);
    while (tile > GBMV_NT && 2 * (s1 + s2 + sx) * (int64_t)sizeof(T) + 128 > h->max_smem_optin - 1024) {
        tile -= GBMV_NT;
        stage_elems(tile, s1, s2, sx);
    }
    a.tile = tile; a.ntiles = (m + tile - 1) / tile; a.sA_elems[0] = s1; a.sA_elems[1] = s2; a.sx_elems = sx;
    const int64_t stage_bytes = (s1 + s2 + sx) * (int64_t)sizeof(T);
    int64_t stages = h->tune.gbmv_stages > 0 ? h->tune.gbmv_stages : 2;
    if (stages > GBMV_MAX_STAGES) stages = GBMV_MAX_STAGES;
    int64_t ctas = h->tune.gbmv_ctas_per_sm > 0 ? h->tune.gbmv_ctas_per_sm : (227 * 1024) / (stages * stage_bytes + 128 + 1024);
    if (ctas < 1) ctas = 1;
    if (ctas > 8) ctas = 8;
    while (stages > 1 && ctas * (stages * stage_bytes + 128 + 1024) > 227 * 1024) --stages;
    while (ctas > 1 && ctas * (stages * stage_bytes + 128 + 1024) > 227 * 1024) --ctas;
    a.stages = (int)stages;
    int64_t grid = (int64_t)h->sm_count * ctas;
    if (grid > a.ntiles) grid = a.ntiles;
    const size_t smem = (size_t)(128 + stages * stage_bytes);
    const int64_t nbm = nb1 > nb2 ? nb1 : nb2;
    if (nbm <= 3) return launch_gbmv2<T, 3>(h, a, (int)grid, smem);
    if (nbm <= 5) return launch_gbmv2<T, 5>(h, a, (int)grid, smem);
    if (nbm <= 9) return launch_gbmv2<T, 9>(h, a, (int)grid, smem);
    if (nbm <= 17) return launch_gbmv2<T, 17>(h, a, (int)grid, smem);
    return launch_gbmv2<T, 33>(h, a, (int)grid, smem);
}

// Host-buffer entry point: H2D(A, x) + kernel + D2H(y), pipelined.  op = N streams A in column
// chunks on a copy stream; as soon as the columns a block of rows needs have landed, that block's
// sub-slab (again a rectangular banded matrix: same lda, shifted bandwidths) is applied on the
// compute stream and its piece of y goes back on a third stream -- so kernels and the D2H hide under
// the H2D, which is the PCIe-bound critical path.
template <typename T>
int gbmv_host_impl(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, T alpha, const T *A,
                   int64_t lda, const T *x, T beta, T *y) {
    if (!h) return -1;
    const bool tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!tr && !(trans == 'N' || trans == 'n')) return lbm_fail_arg(h, 2, "trans must be N, T or C");
    if (m < 0) return lbm_fail_arg(h, 3, "m < 0");
    if (n < 0) return lbm_fail_arg(h, 4, "n < 0");
    if (kl < 0) return lbm_fail_arg(h, 5, "kl < 0");
    if (ku < 0) return lbm_fail_arg(h, 6, "ku < 0");
    if (lda < kl + ku + 1) return lbm_fail_arg(h, 9, "lda < kl+ku+1");
    const int64_t nout = tr ? n : m, nin = tr ? m : n;
    if (nout == 0) return 0;
    if (!y) return lbm_fail_arg(h, 12, "y == NULL");
    if (nin > 0 && !A) return lbm_fail_arg(h, 8, "A == NULL");
    if (nin > 0 && !x) return lbm_fail_arg(h, 10, "x == NULL");
    lbm_device_guard g(h->device);
    void *dAv, *dxv, *dyv;
    int rc;
    if ((rc = lbm_stage(h, 0, (size_t)(lda * n) * sizeof(T) + 16, &dAv))) return rc;
    if ((rc = lbm_stage(h, 1, (size_t)nin * sizeof(T) + 16, &dxv))) return rc;
    if ((rc = lbm_stage(h, 2, (size_t)nout * sizeof(T) + 16, &dyv))) return rc;
    T *dA = (T *)dAv, *dx = (T *)dxv, *dy = (T *)dyv;
    const int64_t chunk_cols = ((int64_t)256 << 20) / (lda * (int64_t)sizeof(T));  // ~256 MB of A per chunk
    const bool pipelined = !tr && n > 2 * chunk_cols && chunk_cols > 4 * (kl + ku + 1);
    if (!pipelined) {
        if (lda * n > 0) LBM_CUDA(h, cudaMemcpyAsync(dA, A, (size_t)(lda * n) * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        if (nin) LBM_CUDA(h, cudaMemcpyAsync(dx, x, (size_t)nin * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        if (beta != (T)0) LBM_CUDA(h, cudaMemcpyAsync(dy, y, (size_t)nout * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        rc = gbmv_impl<T>(h, trans, m, n, kl, ku, alpha, dA, lda, dx, beta, dy);
        if (rc) {  // the header's promise: `_host` calls return with no DMA left in flight on the caller's buffers
            cudaStreamSynchronize(h->stream);
            return rc;
        }
        LBM_CUDA(h, cudaMemcpyAsync(y, dy, (size_t)nout * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
        LBM_CUDA(h, cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (!h->h2d_stream) {
        LBM_CUDA(h, cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
        LBM_CUDA(h, cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
        LBM_CUDA(h, cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
        LBM_CUDA(h, cudaEventCreateWithFlags(&h->ev_kern, cudaEventDisableTiming));
    }
    cudaStream_t sc = h->h2d_stream, sk = h->stream, so = h->d2h_stream;
    // the pipeline proper; ANY failure inside still drains the three streams before returning, so the caller may free or
    // reuse A, x and y (the header's promise for `_host` entry points)
    auto pipeline = [&]() -> int {
        // order the pipeline after whatever is already queued on the handle's stream
        LBM_CUDA(h, cudaEventRecord(h->ev_kern, sk));
        LBM_CUDA(h, cudaStreamWaitEvent(sc, h->ev_kern, 0));
        LBM_CUDA(h, cudaStreamWaitEvent(so, h->ev_kern, 0));
        LBM_CUDA(h, cudaMemcpyAsync(dx, x, (size_t)nin * sizeof(T), cudaMemcpyHostToDevice, sc));
        if (beta != (T)0) LBM_CUDA(h, cudaMemcpyAsync(dy, y, (size_t)nout * sizeof(T), cudaMemcpyHostToDevice, sc));
        int64_t r_done = 0;
        for (int64_t c0 = 0; c0 < n; c0 += chunk_cols) {
            const int64_t c1 = c0 + chunk_cols < n ? c0 + chunk_cols : n;
            LBM_CUDA(h, cudaMemcpyAsync(dA + c0 * lda, A + c0 * lda, (size_t)((c1 - c0) * lda) * sizeof(T),
                                        cudaMemcpyHostToDevice, sc));
            LBM_CUDA(h, cudaEventRecord(h->ev_copy, sc));
            // rows whose columns [i-kl, i+ku] all lie in [0, c1): i < c1 - ku (all remaining rows once c1 == n)
            int64_t r_hi = c1 == n ? m : c1 - ku;
            r_hi &= ~(int64_t)3;                   // keep sub-slabs 16-byte aligned
            if (c1 == n) r_hi = m;
            if (r_hi > m) r_hi = m;
            if (r_hi <= r_done) continue;
            LBM_CUDA(h, cudaStreamWaitEvent(sk, h->ev_copy, 0));
            const int64_t a0 = r_done, b0 = r_hi;
            const int64_t cc0 = a0 - kl > 0 ? a0 - kl : 0;
            const int64_t cc1 = b0 + ku < n ? b0 + ku : n;
            const int64_t shift = a0 - cc0;
            int rc2 = gbmv_impl<T>(h, 'N', b0 - a0, cc1 - cc0, kl - shift, ku + shift, alpha, dA + cc0 * lda, lda, dx + cc0,
                                   beta, dy + a0);
            if (rc2) return rc2;
            LBM_CUDA(h, cudaEventRecord(h->ev_kern, sk));
            LBM_CUDA(h, cudaStreamWaitEvent(so, h->ev_kern, 0));
            LBM_CUDA(h, cudaMemcpyAsync(y + a0, dy + a0, (size_t)(b0 - a0) * sizeof(T), cudaMemcpyDeviceToHost, so));
            r_done = r_hi;
        }
        return 0;
    };
    rc = pipeline();
    const cudaError_t e1 = cudaStreamSynchronize(sc), e2 = cudaStreamSynchronize(sk), e3 = cudaStreamSynchronize(so);
    if (rc) return rc;
    if (e1 != cudaSuccess) return lbm_fail_cuda(h, e1, "cudaStreamSynchronize(h2d)");
    if (e2 != cudaSuccess) return lbm_fail_cuda(h, e2, "cudaStreamSynchronize(compute)");
    if (e3 != cudaSuccess) return lbm_fail_cuda(h, e3, "cudaStreamSynchronize(d2h)");
    return 0;
}

}  // namespace


// y <- alpha * F_0 (F_1 ( ... F_{nf-1} x)) + beta * y: the flattened right-to-left apply of a lazy product
// (LazyArrays' Mul flattening, the rule the reference relies on at src/blockkron.jl:208-214) in ONE call: nf launches on
// the handle's stream, the nf-1 temporaries ping-pong inside the handle's scratch buffer.
template <typename T>
int gbmv_chain_impl(lbm_handle_t h, int nf, const int64_t *m, const int64_t *n, const int64_t *kl, const int64_t *ku,
                    const T *const *A, const int64_t *lda, T alpha, const T *x, T beta, T *y) {
    if (!h) return -1;
    if (nf < 1) return lbm_fail_arg(h, 2, "nf < 1");
    if (!m || !n || !kl || !ku || !A || !lda) return lbm_fail_arg(h, 3, "NULL factor description");
    int64_t tmax = 0;
    for (int f = 0; f < nf; ++f) {
        if (m[f] < 0 || n[f] < 0) return lbm_fail_arg(h, 3, "negative factor size");
        if (f + 1 < nf && n[f] != m[f + 1]) return lbm_fail_arg(h, 4, "inner sizes of consecutive factors differ");
        if (f > 0 && m[f] > tmax) tmax = m[f];
    }
    if (nf == 1) return gbmv_impl<T>(h, 'N', m[0], n[0], kl[0], ku[0], alpha, A[0], lda[0], x, beta, y);
    const int64_t pitch = round_up(tmax > 0 ? tmax : 1, 32);
    void *scr;
    int rc = lbm_scratch(h, (size_t)(2 * pitch) * sizeof(T), &scr);
    if (rc) return rc;
    T *tmp[2] = {(T *)scr, (T *)scr + pitch};
    const T *v = x;
    for (int f = nf - 1; f >= 1; --f) {
        T *t = tmp[f & 1];
        rc = gbmv_impl<T>(h, 'N', m[f], n[f], kl[f], ku[f], (T)1, A[f], lda[f], v, (T)0, t);
        if (rc) return rc;
        v = t;
    }
    return gbmv_impl<T>(h, 'N', m[0], n[0], kl[0], ku[0], alpha, A[0], lda[0], v, beta, y);
}

// dist.cu: the row-sharded apply with the ghost exchange fused into the launch (returns LBM_NOT_FUSED when the
// streamed-tile kernel does not cover the shape; nothing has been launched then)
int lbm_internal_dgbmv_fused(lbm_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, double alpha, const double *A,
                             int64_t lda, const double *x, double beta, double *y, const lbm::HaloFuse *hf) {
    return gbmv_impl<double>(h, 'N', m, n, kl, ku, alpha, A, lda, x, beta, y, hf);
}
int lbm_internal_sgbmv_fused(lbm_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, float alpha, const float *A,
                             int64_t lda, const float *x, float beta, float *y, const lbm::HaloFuse *hf) {
    return gbmv_impl<float>(h, 'N', m, n, kl, ku, alpha, A, lda, x, beta, y, hf);
}

extern "C" {
int lbm_dgbmv_chain(lbm_handle_t h, int nf, const int64_t *m, const int64_t *n, const int64_t *kl, const int64_t *ku,
                    const double *const *A, const int64_t *lda, double alpha, const double *x, double beta, double *y) {
    return gbmv_chain_impl<double>(h, nf, m, n, kl, ku, A, lda, alpha, x, beta, y);
}
int lbm_sgbmv_chain(lbm_handle_t h, int nf, const int64_t *m, const int64_t *n, const int64_t *kl, const int64_t *ku,
                    const float *const *A, const int64_t *lda, float alpha, const float *x, float beta, float *y) {
    return gbmv_chain_impl<float>(h, nf, m, n, kl, ku, A, lda, alpha, x, beta, y);
}
int lbm_dgbmv(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, double alpha,
              const double *A, int64_t lda, const double *x, double beta, double *y) {
    return gbmv_impl<double>(h, trans, m, n, kl, ku, alpha, A, lda, x, beta, y);
}
int lbm_sgbmv(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, float alpha, const float *A,
              int64_t lda, const float *x, float beta, float *y) {
    return gbmv_impl<float>(h, trans, m, n, kl, ku, alpha, A, lda, x, beta, y);
}
int lbm_dgbmv_multi(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, double alpha,
                    const double *A, int64_t lda, const double *X, int64_t ldx, double beta, double *Y, int64_t ldy,
                    int64_t nrhs) {
    return gbmv_multi_impl<double>(h, trans, m, n, kl, ku, alpha, A, lda, X, ldx, beta, Y, ldy, nrhs);
}
int lbm_sgbmv_multi(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, float alpha,
                    const float *A, int64_t lda, const float *X, int64_t ldx, float beta, float *Y, int64_t ldy,
                    int64_t nrhs) {
    return gbmv_multi_impl<float>(h, trans, m, n, kl, ku, alpha, A, lda, X, ldx, beta, Y, ldy, nrhs);
}
int lbm_dgbmv2(lbm_handle_t h, int64_t m, int64_t n, double alpha1, int64_t kl1, int64_t ku1, const double *A1, int64_t lda1,
               double alpha2, int64_t kl2, int64_t ku2, const double *A2, int64_t lda2, const double *x, double beta,
               double *y) {
    return gbmv2_impl<double>(h, m, n, alpha1, kl1, ku1, A1, lda1, alpha2, kl2, ku2, A2, lda2, x, beta, y);
}
int lbm_sgbmv2(lbm_handle_t h, int64_t m, int64_t n, float alpha1, int64_t kl1, int64_t ku1, const float *A1, int64_t lda1,
               float alpha2, int64_t kl2, int64_t ku2, const float *A2, int64_t lda2, const float *x, float beta, float *y) {
    return gbmv2_impl<float>(h, m, n, alpha1, kl1, ku1, A1, lda1, alpha2, kl2, ku2, A2, lda2, x, beta, y);
}
int lbm_dgbmv_host(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, double alpha,
                   const double *A, int64_t lda, const double *x, double beta, double *y) {
    return gbmv_host_impl<double>(h, trans, m, n, kl, ku, alpha, A, lda, x, beta, y);
}
int lbm_sgbmv_host(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, float alpha,
                   const float *A, int64_t lda, const float *x, float beta, float *y) {
    return gbmv_host_impl<float>(h, trans, m, n, kl, ku, alpha, A, lda, x, beta, y);
}
}
