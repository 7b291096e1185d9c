// gbmm_reg.cu -- narrow banded x banded -> banded with the output band kept in registers.
//
// Materialises ApplyMatrix(*, A, B) (reference README.md:8-26; upstream BandedMatrices gbmm! = n tiny gbmv ccalls,
// SURVEY.md 3.3) for bands up to 17 stored diagonals per factor -- the l,u <= 8 regime of the north-star target,
// which is HBM-bound: bytes = w*n*[(la+ua+1) + (lb+ub+1) + (la+lb+ua+ub+1)].
//
// In diagonal-major storage the product is a convolution of band rows:
//     Cdata[r + s, j] += Adata[r, j + s - kub] * Bdata[s, j],   r in [0, NA), s in [0, NB)     (kuc = kua + kub)
// so only the STORED ROW COUNTS NA = kla+kua+1, NB = klb+kub+1 shape the loop nest, never the individual (kl,ku).
// A thread owns two adjacent output columns and their NA+NB-1 accumulators each, all in registers (compile-time
// indices: the loops below are fully unrolled).  The slab column of A at local index 2t+q feeds column 2t with s = q
// and column 2t+1 with s = q-1, so every value of A read from shared memory is used twice and shared-memory
// bandwidth (the limiter of a one-column-per-thread version: one LDS per DFMA) stays below the HBM time.
//
// Staging: the slab of A (columns [j0-kub, j0+TC+klb), ONE contiguous range), the tile of B and, on the way out, the
// tile of C pass through shared memory so every global access is a flat, fully coalesced stream.  Shared-memory
// columns are laid out in PAIRS with an odd pair pitch (2*P+1 words): lane t reads column 2t+q, i.e. lanes are one
// pair apart, and an odd pitch makes those LDS.64 / STS.64 conflict-free.  Template row counts NAT >= NA, NBT >= NB come
// from {3,5,9,17}; missing rows are zero-filled at staging time, which also masks, BY INDEX, every out-of-matrix slot
// (brand() fills them with random numbers; undef storage may hold NaN).
#include "lbm_common.cuh"

namespace {

constexpr int GR_NT = 128;          // threads per CTA
constexpr int GR_TC = 2 * GR_NT;    // output columns per CTA

template <typename T>
struct GrArgs {
    int64_t m, n, k;
    int kla, kua, klb, kub, klc, kuc;
    int64_t lda, ldb, ldc;
    T alpha, beta;
    const T *A;
    const T *B;
    T *C;
};

// resident CTAs per SM the register budget is tuned for: the wide instantiations hold 2*(NAT+NBT-1) accumulators
constexpr int gr_min_blocks(int nat, int nbt) { return nat + nbt >= 26 ? 3 : (nat + nbt >= 14 ? 5 : (nat + nbt >= 8 ? 8 : 12)); }

// one element global -> shared, asynchronously (LDGSTS); ok == false writes a zero without touching global memory
template <typename T>
__device__ __forceinline__ void cp_async_elem(uint32_t dst, const T *src, bool ok) {
    const int nbytes = ok ? (int)sizeof(T) : 0;
    if (sizeof(T) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}

// pair-skewed address of (column c, row r) for row pitch P
template <int P>
__device__ __forceinline__ int skew(int c, int r) {
    return (c >> 1) * (2 * P + 1) + (c & 1) * P + r;
}

template <typename T, int NAT, int NBT>
__global__ void __launch_bounds__(GR_NT, gr_min_blocks(NAT, NBT)) gbmm_reg_kernel(const GrArgs<T> a) {
    constexpr int NCT = NAT + NBT - 1;
    constexpr int SA_COLS = GR_TC + NBT + 1;                       // slab columns (even count)
    constexpr int SA_ELEMS = (SA_COLS / 2) * (2 * NAT + 1);
    constexpr int SB_ELEMS = (GR_TC / 2) * (2 * NBT + 1);
    constexpr int SC_ELEMS = (GR_TC / 2) * (2 * NCT + 1);
    constexpr int SM_ELEMS = (SA_ELEMS + SB_ELEMS) > SC_ELEMS ? (SA_ELEMS + SB_ELEMS) : SC_ELEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    T *sB = sA + SA_ELEMS;
    T *sC = sA;                                                    // aliases sA/sB after the compute phase
    static_assert(SM_ELEMS >= SC_ELEMS, "smem");

    const int tid = threadIdx.x;
    const int NA = a.kla + a.kua + 1, NB = a.klb + a.kub + 1;
    const int64_t ntiles = (a.n + GR_TC - 1) / GR_TC;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t j0 = tile * GR_TC;
        const int cols = (int)((a.n - j0) < GR_TC ? (a.n - j0) : GR_TC);
        const int64_t p0 = j0 - a.kub;                             // global column of slab column 0
        // interior tile: every staged slot is inside the matrix, no index tests needed
        const bool interior = cols == GR_TC && p0 - a.kua >= 0 && p0 + SA_COLS - 1 + a.kla < a.m && p0 >= 0 &&
                              p0 + SA_COLS - 1 < a.k && j0 - a.kub >= 0 && j0 + GR_TC - 1 + a.klb < a.k;

        // ---- stage A slab (pitch NAT, rows >= NA zero) and B tile (pitch NBT) ----
        // element-wise cp.async (LDGSTS) straight into the skewed layout: nothing waits on a load inside the loop, so
        // the whole tile (~70 elements per thread) is in flight at once; src-size 0 zero-fills masked / padding slots
        {
            const T *Ag = a.A + p0 * a.lda;                        // may point before A for p0 < 0: never dereferenced there
            const uint32_t sA_u = lbm::smem_u32(sA), sB_u = lbm::smem_u32(sB);
            // skew<P>(c, r) = c*P + r + (c >> 1) = e + (c >> 1): with a tight leading dimension (lda == NAT) the global
            // offset is e itself, so an interior element costs one multiply-high, one shift and the copy
            if (interior && a.lda == NAT) {
#pragma unroll 4
                for (int e = tid; e < SA_COLS * NAT; e += GR_NT)
                    cp_async_elem<T>(sA_u + (uint32_t)((e + ((e / NAT) >> 1)) * (int)sizeof(T)), Ag + e, true);
            } else {
                for (int e = tid; e < SA_COLS * NAT; e += GR_NT) {
                    const int c = e / NAT, r = e - c * NAT;
                    bool ok = r < NA;
                    if (ok && !interior) {
                        const int64_t p = p0 + c, i = p + r - a.kua;
                        ok = p >= 0 && p < a.k && i >= 0 && i < a.m;
                    }
                    cp_async_elem<T>(sA_u + (uint32_t)(skew<NAT>(c, r) * (int)sizeof(T)), ok ? Ag + (r + a.lda * c) : a.A, ok);
                }
            }
            const T *Bg = a.B + j0 * a.ldb;
            if (interior && a.ldb == NBT) {
#pragma unroll 4
                for (int e = tid; e < GR_TC * NBT; e += GR_NT)
                    cp_async_elem<T>(sB_u + (uint32_t)((e + ((e / NBT) >> 1)) * (int)sizeof(T)), Bg + e, true);
            } else {
                for (int e = tid; e < GR_TC * NBT; e += GR_NT) {
                    const int c = e / NBT, s = e - c * NBT;
                    bool ok = s < NB && c < cols;
                    if (ok && !interior) {
                        const int64_t p = j0 + c + s - a.kub;
                        ok = p >= 0 && p < a.k;
                    }
                    cp_async_elem<T>(sB_u + (uint32_t)(skew<NBT>(c, s) * (int)sizeof(T)), ok ? Bg + (s + a.ldb * c) : a.B, ok);
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();

        // ---- compute: two columns, NCT accumulators each, everything in registers ----
        T acc0[NCT], acc1[NCT];
#pragma unroll
        for (int t = 0; t < NCT; ++t) acc0[t] = acc1[t] = (T)0;
        {
            const T *pa = sA + tid * (2 * NAT + 1);                // pair `tid`; column 2*tid + q lives q/2 pairs further on
            const T *pb = sB + tid * (2 * NBT + 1);
#pragma unroll
            for (int q = 0; q <= NBT; ++q) {
                T av[NAT];
                const T *col = pa + (q >> 1) * (2 * NAT + 1) + (q & 1) * NAT;
#pragma unroll
                for (int r = 0; r < NAT; ++r) av[r] = col[r];
                if (q < NBT) {
                    const T b0 = pb[q];
#pragma unroll
                    for (int r = 0; r < NAT; ++r) acc0[r + q] = fma(av[r], b0, acc0[r + q]);
                }
                if (q >= 1) {
                    const T b1 = pb[NBT + q - 1];
#pragma unroll
                    for (int r = 0; r < NAT; ++r) acc1[r + q - 1] = fma(av[r], b1, acc1[r + q - 1]);
                }
            }
        }
        __syncthreads();                                           // all reads of sA / sB done: sC may overwrite them

        // ---- stage the C tile (pair-skewed, pitch NCT) and stream it out ----
        {
            T *pc = sC + tid * (2 * NCT + 1);
#pragma unroll
            for (int t = 0; t < NCT; ++t) {
                pc[t] = acc0[t];
                pc[NCT + t] = acc1[t];
            }
        }
        __syncthreads();
        {
            const int nbc = a.klc + a.kuc + 1;
            const int off = a.kuc - a.kua - a.kub;                 // C band row of accumulator 0
            const int NC = NA + NB - 1;
            T *Cg = a.C + j0 * a.ldc;
            const int total = cols * nbc;
            const bool fast_out = nbc == NCT && a.ldc == NCT && off == 0 && a.beta == (T)0 && cols == GR_TC &&
                                  j0 - a.kuc >= 0 && j0 + GR_TC - 1 + a.klc < a.m;
            if (fast_out) {
                // tight, fully in-matrix tile: global offset = e, shared index = e + (c >> 1)
#pragma unroll 4
                for (int e = tid; e < GR_TC * NCT; e += GR_NT) Cg[e] = a.alpha * sC[e + ((e / NCT) >> 1)];
            } else {
                // flat element index e = c*nbc + row, advanced incrementally (no division in the loop)
                const int dc = GR_NT / nbc, dr = GR_NT - dc * nbc;
                int c = tid / nbc, row = tid - c * nbc;
                for (int e = tid; e < total; e += GR_NT) {
                    const int t = row - off;
                    const int64_t i = j0 + c + row - a.kuc;
                    const bool in = i >= 0 && i < a.m;
                    T v = (T)0;
                    if (t >= 0 && t < NC && in) v = sC[skew<NCT>(c, t)];
                    T *dst = Cg + row + a.ldc * c;
                    if (a.beta == (T)0) *dst = a.alpha * v;
                    else if (in) *dst = fma(a.beta, *dst, a.alpha * v);
                    row += dr;
                    c += dc;
                    if (row >= nbc) { row -= nbc; ++c; }
                }
            }
        }
        __syncthreads();                                           // sC aliases the next tile's sA / sB
    }
}

template <int NAT, int NBT, typename T>
constexpr size_t gr_smem_bytes() {
    constexpr int NCT = NAT + NBT - 1;
    constexpr int SA_COLS = GR_TC + NBT + 1;
    constexpr int SA_ELEMS = (SA_COLS / 2) * (2 * NAT + 1);
    constexpr int SB_ELEMS = (GR_TC / 2) * (2 * NBT + 1);
    constexpr int SC_ELEMS = (GR_TC / 2) * (2 * NCT + 1);
    return sizeof(T) * (size_t)((SA_ELEMS + SB_ELEMS) > SC_ELEMS ? (SA_ELEMS + SB_ELEMS) : SC_ELEMS);
}

template <typename T, int NAT, int NBT>
int gr_launch(lbm_handle_t h, const GrArgs<T> &a) {
    static_assert((GR_TC + NBT + 1) % 2 == 0, "slab column count must be even");
    constexpr size_t smem = gr_smem_bytes<NAT, NBT, T>();
    auto kern = gbmm_reg_kernel<T, NAT, NBT>;
    static int configured = 0;
    if (!(configured & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured |= (1 << h->device);
    }
    int per_sm = 1;
    LBM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GR_NT, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (a.n + GR_TC - 1) / GR_TC;
    int64_t grid = (int64_t)h->sm_count * per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, GR_NT, smem, h->stream>>>(a);
    LBM_LAUNCH_CHECK(h, "gbmm_reg_kernel");
    return 0;
}

template <typename T, int NAT>
int gr_dispatch_b(lbm_handle_t h, const GrArgs<T> &a, int NB) {
    if (NB <= 3) return gr_launch<T, NAT, 3>(h, a);
    if (NB <= 5) return gr_launch<T, NAT, 5>(h, a);
    if (NB <= 9) return gr_launch<T, NAT, 9>(h, a);
    return gr_launch<T, NAT, 17>(h, a);
}

}  // namespace

// Returns 0 and sets *handled when the register kernel covers the shape (both factors <= 17 stored diagonals).
template <typename T>
int lbm_gbmm_reg(lbm_handle_t h, int64_t m, int64_t n, int64_t k, T alpha, int64_t kla, int64_t kua, const T *A, int64_t lda,
                 int64_t klb, int64_t kub, const T *B, int64_t ldb, T beta, int64_t klc, int64_t kuc, T *C, int64_t ldc,
                 bool *handled) {
    *handled = false;
    const int64_t NA = kla + kua + 1, NB = klb + kub + 1;
    if (NA > 17 || NB > 17 || m <= 0 || n <= 0 || k <= 0) return 0;
    if (kuc < kua + kub) return 0;       // C's band clipped by the matrix size (tiny matrices): generic kernel
    if (klc + kuc + 1 > 4096 || lda > (1 << 20) || ldb > (1 << 20) || ldc > (1 << 20)) return 0;
    *handled = true;
    GrArgs<T> a;
    a.m = m; a.n = n; a.k = k;
    a.kla = (int)kla; a.kua = (int)kua; a.klb = (int)klb; a.kub = (int)kub; a.klc = (int)klc; a.kuc = (int)kuc;
    a.lda = lda; a.ldb = ldb; a.ldc = ldc; a.alpha = alpha; a.beta = beta; a.A = A; a.B = B; a.C = C;
    if (NA <= 3) return gr_dispatch_b<T, 3>(h, a, (int)NB);
    if (NA <= 5) return gr_dispatch_b<T, 5>(h, a, (int)NB);
    if (NA <= 9) return gr_dispatch_b<T, 9>(h, a, (int)NB);
    return gr_dispatch_b<T, 17>(h, a, (int)NB);
}

template int lbm_gbmm_reg<double>(lbm_handle_t, int64_t, int64_t, int64_t, double, int64_t, int64_t, const double *, int64_t,
                                  int64_t, int64_t, const double *, int64_t, double, int64_t, int64_t, double *, int64_t, bool *);
template int lbm_gbmm_reg<float>(lbm_handle_t, int64_t, int64_t, int64_t, float, int64_t, int64_t, const float *, int64_t,
                                 int64_t, int64_t, const float *, int64_t, float, int64_t, int64_t, float *, int64_t, bool *);
