// gbmm.cu -- banded x banded -> banded:  C <- alpha*A*B + beta*C, all three in diagonal-major
// storage.  Materialises ApplyMatrix(*, A, B) (reference README.md:8-26); upstream
// BandedMatrices gbmm! does this as n tiny gbmv ccalls.
//
//   C[i,j] = sum_p A[i,p] B[p,j],  p in [max(i-kla, j-kub, 0), min(i+kua, j+klb, k-1)]
//
//  * gbmm_diag_kernel: "diagonal-wise": a CTA owns a tile of output columns; the slab of A it
//    needs (columns [j0-kub, j1+klb), one contiguous range) and the tile of B are staged in
//    shared memory; consecutive threads own consecutive band rows of one output column, so
//    A is read at consecutive smem addresses, B as a broadcast, and C is written contiguously.
//  * gbmm_chain3_kernel: D = (A*B)*C for a batch of small-band factors with the intermediate
//    band A*B living only in shared memory.
//  * wide bands (FP64 DMMA tiles) live in gbmm_dmma.cu.
#include "lbm_common.cuh"

int lbm_chain3_warp_f32(lbm_handle_t h, int64_t batch, int64_t n, const int64_t *kl, const int64_t *ku, const float *A,
                        int64_t sA, const float *B, int64_t sB, const float *C, int64_t sC, float *D, int64_t sD,
                        bool *handled);
int lbm_gbmm_dmma_f64(lbm_handle_t h, int64_t m, int64_t n, int64_t k, double alpha, int64_t kla, int64_t kua,
                      const double *A, int64_t lda, int64_t klb, int64_t kub, const double *B, int64_t ldb,
                      double beta, int64_t klc, int64_t kuc, double *C, int64_t ldc, bool *handled);

// narrow bands (<= 17 stored diagonals per factor): output band in registers, gbmm_reg.cu
template <typename T>
int lbm_gbmm_reg(lbm_handle_t h, int64_t m, int64_t n, int64_t k, T alpha, int64_t kla, int64_t kua, const T *A, int64_t lda,
                 int64_t klb, int64_t kub, const T *B, int64_t ldb, T beta, int64_t klc, int64_t kuc, T *C, int64_t ldc,
                 bool *handled);

namespace {

using namespace lbm;

template <typename T>
struct GbmmArgs {
    int64_t m, n, k;
    int64_t kla, kua, lda, klb, kub, ldb, klc, kuc, ldc;
    T alpha, beta;
    const T *A;
    const T *B;
    T *C;
    int tile;  // output columns per CTA
    int64_t sA_elems;
};

template <typename T, bool USE_SMEM>
__global__ void __launch_bounds__(256) gbmm_diag_kernel(const GbmmArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    T *sB = sA + a.sA_elems;
    const int tid = threadIdx.x;
    const int64_t j0 = (int64_t)blockIdx.x * a.tile;
    const int64_t j1 = j0 + a.tile < a.n ? j0 + a.tile : a.n;
    int64_t p0 = j0 - a.kub > 0 ? j0 - a.kub : 0;       // A columns needed: [p0, p1)
    int64_t p1 = j1 + a.klb < a.k ? j1 + a.klb : a.k;
    if (p1 < p0) p1 = p0;
    const T *Ag = a.A + p0 * a.lda;
    const T *Bg = a.B + j0 * a.ldb;
    if (USE_SMEM) {
        const int64_t cntA = (p1 - p0) * a.lda;
        const int64_t cntB = (j1 - j0) * a.ldb;
        for (int64_t e = tid; e < cntA; e += 256) sA[e] = Ag[e];
        for (int64_t e = tid; e < cntB; e += 256) sB[e] = Bg[e];
        __syncthreads();
    }
    const T *Ar = USE_SMEM ? sA : Ag;
    const T *Br = USE_SMEM ? sB : Bg;
    const int nbc = (int)(a.klc + a.kuc + 1);
    const int64_t total = (j1 - j0) * nbc;
    for (int64_t e = tid; e < total; e += 256) {
        const int jj = (int)(e / nbc);
        const int r = (int)(e - (int64_t)jj * nbc);
        const int64_t j = j0 + jj;
        const int64_t i = j + r - a.kuc;
        T *c = a.C + r + a.ldc * j;
        if (i < 0 || i >= a.m) {
            if (a.beta == (T)0) *c = (T)0;
            continue;
        }
        int64_t plo = i - a.kla > j - a.kub ? i - a.kla : j - a.kub;
        int64_t phi = i + a.kua < j + a.klb ? i + a.kua : j + a.klb;
        if (plo < 0) plo = 0;
        if (phi > a.k - 1) phi = a.k - 1;
        T acc = 0;
        // A index: (kua + i - p) + lda*(p - p0);  B index: (kub + p - j) + ldb*jj
        int64_t ia = (a.kua + i - plo) + a.lda * (plo - p0);
        int64_t ib = (a.kub + plo - j) + a.ldb * jj;
        for (int64_t p = plo; p <= phi; ++p) {
            acc = fma(Ar[ia], Br[ib], acc);
            ia += a.lda - 1;
            ib += 1;
        }
        *c = (a.beta == (T)0) ? a.alpha * acc : fma(a.beta, *c, a.alpha * acc);
    }
}

template <typename T>
int gbmm_impl(lbm_handle_t h, int64_t m, int64_t n, int64_t k, T alpha, int64_t kla, int64_t kua, const T *A,
              int64_t lda, int64_t klb, int64_t kub, const T *B, int64_t ldb, T beta, int64_t klc, int64_t kuc, T *C,
              int64_t ldc) {
    if (!h) return -1;
    if (m < 0) return lbm_fail_arg(h, 2, "m < 0");
    if (n < 0) return lbm_fail_arg(h, 3, "n < 0");
    if (k < 0) return lbm_fail_arg(h, 4, "k < 0");
    if (kla < 0) return lbm_fail_arg(h, 6, "kla < 0");
    if (kua < 0) return lbm_fail_arg(h, 7, "kua < 0");
    if (lda < kla + kua + 1) return lbm_fail_arg(h, 9, "lda < kla+kua+1");
    if (klb < 0) return lbm_fail_arg(h, 10, "klb < 0");
    if (kub < 0) return lbm_fail_arg(h, 11, "kub < 0");
    if (ldb < klb + kub + 1) return lbm_fail_arg(h, 13, "ldb < klb+kub+1");
    if (klc < 0) return lbm_fail_arg(h, 15, "klc < 0");
    if (kuc < 0) return lbm_fail_arg(h, 16, "kuc < 0");
    if (ldc < klc + kuc + 1) return lbm_fail_arg(h, 18, "ldc < klc+kuc+1");
    // C's band must hold every structurally non-zero entry of the product
    const int64_t need_l = kla + klb < m - 1 ? kla + klb : m - 1;
    const int64_t need_u = kua + kub < n - 1 ? kua + kub : n - 1;
    if (k > 0 && klc < need_l) return lbm_fail_arg(h, 15, "klc < min(kla+klb, m-1)");
    if (k > 0 && kuc < need_u) return lbm_fail_arg(h, 16, "kuc < min(kua+kub, n-1)");
    if (n == 0) return 0;
    if (!C) return lbm_fail_arg(h, 17, "C == NULL");
    if (k > 0 && !A) return lbm_fail_arg(h, 8, "A == NULL");
    if (k > 0 && !B) return lbm_fail_arg(h, 12, "B == NULL");
    lbm_device_guard g(h->device);

    // gbmm_force: 0 auto, 1 generic diagonal-wise kernel, 2 DMMA, 3 register kernel; the register kernel wants enough
    // columns to fill the machine (C1's n = 10^4 stays on the latency-optimised diagonal-wise kernel)
    if (h->tune.gbmm_force == 3 || h->tune.gbmm_force == 4 || (h->tune.gbmm_force == 0 && n >= 256 * (int64_t)h->sm_count)) {
        bool handled = false;
        int rc = lbm_gbmm_reg<T>(h, m, n, k, alpha, kla, kua, A, lda, klb, kub, B, ldb, beta, klc, kuc, C, ldc, &handled);
        if (handled) return rc;
    }

    GbmmArgs<T> a;
    a.m = m; a.n = n; a.k = k; a.kla = kla; a.kua = kua; a.lda = lda; a.klb = klb; a.kub = kub; a.ldb = ldb;
    a.klc = klc; a.kuc = kuc; a.ldc = ldc; a.alpha = alpha; a.beta = beta; a.A = A; a.B = B; a.C = C;
    int64_t tile = h->tune.gbmm_tile > 0 ? h->tune.gbmm_tile : 128;
    // keep >= 2 CTAs per SM worth of tiles when the matrix is small
    while (tile > 16 && (n + tile - 1) / tile < 2 * h->sm_count) tile /= 2;
    const int64_t smem_cap = 96 * 1024;
    auto smem_need = [&](int64_t tl) { return ((tl + klb + kub) * lda + tl * ldb) * (int64_t)sizeof(T); };
    while (tile > 8 && smem_need(tile) > smem_cap) tile /= 2;
    a.tile = (int)tile;
    a.sA_elems = (tile + klb + kub) * lda;
    const int64_t grid = (n + tile - 1) / tile;
    if (smem_need(tile) <= smem_cap) {
        static std::atomic<int> configured_dev_mask{0};
        auto kern = gbmm_diag_kernel<T, true>;
        if (!(configured_dev_mask & (1 << h->device))) {
            LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
            configured_dev_mask |= (1 << h->device);
        }
        kern<<<(unsigned)grid, 256, (size_t)smem_need(tile), h->stream>>>(a);
    } else {
        gbmm_diag_kernel<T, false><<<(unsigned)grid, 256, 0, h->stream>>>(a);
    }
    LBM_LAUNCH_CHECK(h, "gbmm_diag_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------
// batched fused chain  D_b = (A_b*B_b)*C_b
// ------------------------------------------------------------------------------------------
template <typename T>
struct ChainArgs {
    int64_t batch, n;
    int kla, kua, klb, kub, klc, kuc;
    const T *A;
    const T *B;
    const T *C;
    T *D;
    int64_t sA, sB, sC, sD;  // batch strides (elements)
    int tile;
    int offB, offC, offAB;   // smem offsets (elements)
};

template <typename T>
__global__ void __launch_bounds__(256) gbmm_chain3_kernel(const ChainArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    T *sB = sA + a.offB;
    T *sC = sA + a.offC;
    T *sAB = sA + a.offAB;
    const int tid = threadIdx.x;
    const int64_t n = a.n;
    const int64_t b = blockIdx.y;
    const int lda = a.kla + a.kua + 1, ldb = a.klb + a.kub + 1, ldc = a.klc + a.kuc + 1;
    const int l01 = a.kla + a.klb, u01 = a.kua + a.kub, ld01 = l01 + u01 + 1;
    const int ld3 = l01 + a.klc, ud3 = u01 + a.kuc, ldd = ld3 + ud3 + 1;
    const int64_t j0 = (int64_t)blockIdx.x * a.tile;
    const int64_t j1 = j0 + a.tile < n ? j0 + a.tile : n;
    const int64_t q0 = j0 - a.kuc > 0 ? j0 - a.kuc : 0;   // columns of A*B (and of B) needed
    const int64_t q1 = j1 + a.klc < n ? j1 + a.klc : n;
    const int64_t p0 = q0 - a.kub > 0 ? q0 - a.kub : 0;   // columns of A needed
    const int64_t p1 = q1 + a.klb < n ? q1 + a.klb : n;
    const T *Ag = a.A + b * a.sA + p0 * lda;
    const T *Bg = a.B + b * a.sB + q0 * ldb;
    const T *Cg = a.C + b * a.sC + j0 * ldc;
    for (int64_t e = tid; e < (p1 - p0) * lda; e += 256) sA[e] = Ag[e];
    for (int64_t e = tid; e < (q1 - q0) * ldb; e += 256) sB[e] = Bg[e];
    for (int64_t e = tid; e < (j1 - j0) * ldc; e += 256) sC[e] = Cg[e];
    __syncthreads();
    // phase 1: AB[:, q0..q1) into shared memory (band (l01,u01))
    const int tot1 = (int)(q1 - q0) * ld01;
    for (int e = tid; e < tot1; e += 256) {
        const int qq = e / ld01;
        const int r = e - qq * ld01;
        const int64_t q = q0 + qq;
        const int64_t i = q + r - u01;
        T acc = 0;
        if (i >= 0 && i < n) {
            int64_t plo = i - a.kla > q - a.kub ? i - a.kla : q - a.kub;
            int64_t phi = i + a.kua < q + a.klb ? i + a.kua : q + a.klb;
            if (plo < 0) plo = 0;
            if (phi > n - 1) phi = n - 1;
            int ia = (int)((a.kua + i - plo) + lda * (plo - p0));
            int ib = (int)((a.kub + plo - q) + ldb * qq);
            for (int64_t p = plo; p <= phi; ++p) {
                acc = fma(sA[ia], sB[ib], acc);
                ia += lda - 1;
                ib += 1;
            }
        }
        sAB[e] = acc;
    }
    __syncthreads();
    // phase 2: D[:, j0..j1) = AB * C
    T *Dg = a.D + b * a.sD;
    const int tot2 = (int)(j1 - j0) * ldd;
    for (int e = tid; e < tot2; e += 256) {
        const int jj = e / ldd;
        const int r = e - jj * ldd;
        const int64_t j = j0 + jj;
        const int64_t i = j + r - ud3;
        T acc = 0;
        if (i >= 0 && i < n) {
            int64_t qlo = i - l01 > j - a.kuc ? i - l01 : j - a.kuc;
            int64_t qhi = i + u01 < j + a.klc ? i + u01 : j + a.klc;
            if (qlo < 0) qlo = 0;
            if (qhi > n - 1) qhi = n - 1;
            int ia = (int)((u01 + i - qlo) + ld01 * (qlo - q0));
            int ic = (int)((a.kuc + qlo - j) + ldc * jj);
            for (int64_t q = qlo; q <= qhi; ++q) {
                acc = fma(sAB[ia], sC[ic], acc);
                ia += ld01 - 1;
                ic += 1;
            }
        }
        Dg[r + (int64_t)ldd * j] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// register-blocked fused chain for compile-time bands (the (4,4)^3 Float32 config C5)
// ------------------------------------------------------------------------------------------
// Thread-per-4-columns: a thread keeps the 4 output columns of a product (band rows in
// registers) and streams the 12 operand columns it needs from shared memory, so each smem
// value feeds up to 4 FMAs (~0.33 LDS per MAC instead of 2).  Operand tiles are stored with a
// mod-4 SPLIT column order -- column c lives in slot (c % 4) * H + c / 4 with an odd column
// pitch -- which makes "thread t reads column 4t + w" a stride-(odd) access: bank-conflict-free.
// Out-of-matrix columns are zero-filled, coefficient loads are masked BY INDEX, and garbage in
// out-of-matrix band slots can only reach out-of-matrix outputs, which are stored as 0.
constexpr int CH_NT = 128;           // phase-2 threads (4 columns each) -> 512 columns per CTA
constexpr int CH_TCOL = 4 * CH_NT;

template <typename T, int KLA, int KUA, int KLB, int KUB, int KLC, int KUC>
struct ChainGeom {
    static constexpr int LDA = KLA + KUA + 1, LDB = KLB + KUB + 1, LDC = KLC + KUC + 1;
    static constexpr int L01 = KLA + KLB, U01 = KUA + KUB, LD01 = L01 + U01 + 1;
    static constexpr int L3 = L01 + KLC, U3 = U01 + KUC, LDD = L3 + U3 + 1;
    // halo columns, rounded to multiples of 4 so every tile start stays 4-aligned
    static constexpr int HQ = ((KLC > KUC ? KLC : KUC) + 3) / 4 * 4;           // AB/B halo each side
    static constexpr int HP = HQ + ((KLB > KUB ? KLB : KUB) + 3) / 4 * 4;      // A halo each side
    static constexpr int NQ = CH_TCOL + 2 * HQ;    // AB / B columns per tile
    static constexpr int NP = CH_TCOL + 2 * HP;    // A columns per tile
    static constexpr int PA = LDA | 1, PB = LDB | 1, PC = LDC | 1, P01 = LD01 | 1, PD = LDD | 1;  // odd pitches
    static constexpr int HA4 = NP / 4, HB4 = NQ / 4, HC4 = CH_TCOL / 4;
    static constexpr int OFF_B = NP * PA, OFF_C = OFF_B + NQ * PB, OFF_AB = OFF_C + CH_TCOL * PC;
    // D staging starts at 0 and may overlap every operand tile: all are dead by then
    static constexpr int SMEM_ELEMS = (OFF_AB + NQ * P01) > (CH_TCOL * PD) ? (OFF_AB + NQ * P01) : (CH_TCOL * PD);
};

template <typename T, int KLA, int KUA, int KLB, int KUB, int KLC, int KUC>
__global__ void __launch_bounds__(CH_NT + 32) gbmm_chain3_reg_kernel(int64_t n, const T *__restrict__ A, int64_t sA_,
                                                                    const T *__restrict__ B, int64_t sB_,
                                                                    const T *__restrict__ C, int64_t sC_,
                                                                    T *__restrict__ D, int64_t sD_) {
    using G = ChainGeom<T, KLA, KUA, KLB, KUB, KLC, KUC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    T *sA = sm, *sB = sm + G::OFF_B, *sC = sm + G::OFF_C, *sAB = sm + G::OFF_AB;
    T *sD = sm;  // D staging; only written after every operand tile is dead
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int64_t b = blockIdx.y;
    const int64_t j0 = (int64_t)blockIdx.x * CH_TCOL;
    const int64_t q0 = j0 - G::HQ;  // first AB / B column of the tile (may be < 0)
    const int64_t p0 = j0 - G::HP;  // first A column of the tile
    const T *Ag = A + b * sA_;
    const T *Bg = B + b * sB_;
    const T *Cg = C + b * sC_;
    T *Dg = D + b * sD_;

    // ---- stage operand tiles (coalesced global reads -> split-column smem), zero outside [0,n)
    // All 128-bit loads of a tile are issued back to back (one latency exposure per operand);
    // tile starts are multiples of 4 columns, so with n % 4 == 0 every vector is 16-byte aligned
    // and lies wholly inside or wholly outside the matrix.
    auto stage = [&](T *dst, const T *src_item, int64_t c_first, int ncols_tile, int ld, int pitch, int h4) {
        const int total = ncols_tile * ld;                       // elements of the tile (multiple of 4)
        const int64_t e_lo = (c_first < 0 ? -c_first : 0) * (int64_t)ld;            // first valid tile element
        const int64_t c_end = c_first + ncols_tile < n ? c_first + ncols_tile : n;
        const int64_t e_hi = (c_end - c_first) * (int64_t)ld;                        // one past the last valid
        const T *src = src_item + c_first * ld;                   // tile element e lives at src[e]
        constexpr int VPT = 8;                                    // vectors per thread per batch
        constexpr int EPVL = 16 / (int)sizeof(T);
        using V = typename vec16<T>::type;
        const bool vec_ok = (n % EPVL == 0) && lbm_aligned16_dev(src_item) && (total % EPVL == 0);
        if (vec_ok) {
            for (int base = 0; base < total / EPVL; base += VPT * nthr) {
                V v[VPT];
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const int vi = base + k * nthr + tid;
                    const int64_t e = (int64_t)vi * EPVL;
                    const bool ok = vi < total / EPVL && e >= e_lo && e + EPVL <= e_hi;
                    if (ok) v[k] = ld_stream(reinterpret_cast<const V *>(src + e));
                    else v[k] = V{};
                }
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const int vi = base + k * nthr + tid;
                    if (vi < total / EPVL) {
                        const T *pv = reinterpret_cast<const T *>(&v[k]);
#pragma unroll
                        for (int w = 0; w < EPVL; ++w) {
                            const int e = vi * EPVL + w;
                            const int lc = e / ld, f = e - lc * ld;
                            dst[((lc & 3) * h4 + (lc >> 2)) * pitch + f] = pv[w];
                        }
                    }
                }
            }
        } else {
            for (int e = tid; e < total; e += nthr) {
                const int lc = e / ld, f = e - lc * ld;
                dst[((lc & 3) * h4 + (lc >> 2)) * pitch + f] = (e >= e_lo && e < e_hi) ? src[e] : (T)0;
            }
        }
    };
    stage(sA, Ag, p0, G::NP, G::LDA, G::PA, G::HA4);
    stage(sB, Bg, q0, G::NQ, G::LDB, G::PB, G::HB4);
    stage(sC, Cg, j0, CH_TCOL, G::LDC, G::PC, G::HC4);
    __syncthreads();

    // ---- phase 1: AB[:, q0 + 4u + w], w = 0..3, for u = tid < NQ/4 --------------------------
    if (tid < G::NQ / 4) {
        const int u = tid;
        T acc[4][G::LD01];
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int r = 0; r < G::LD01; ++r) acc[w][r] = (T)0;
        // B coefficients of the 4 columns: bco[w][e'] = B[q+e, q], e = e' - KUB, masked by index
        T bco[4][G::LDB];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int64_t q = q0 + 4 * u + w;
#pragma unroll
            for (int ep = 0; ep < G::LDB; ++ep) {
                const int64_t p = q + ep - KUB;
                bco[w][ep] = (q >= 0 && q < n && p >= 0 && p < n) ? sB[(w * G::HB4 + u) * G::PB + ep] : (T)0;
            }
        }
        // A columns needed: p = q + e for the 4 q's: local A index = (q0 - p0) + 4u + w + e, i.e.
        // la = 4u + v with v = (HP - HQ) - KUB .. (HP - HQ) + 3 + KLB
        constexpr int V0 = (G::HP - G::HQ) - KUB;
        constexpr int NV = 4 + KLB + KUB;
#pragma unroll
        for (int vv = 0; vv < NV; ++vv) {
            const int v = V0 + vv;                     // column offset relative to 4u (may be negative)
            const int vq = (v % 4 + 4) % 4;            // (4u + v) % 4
            const int vd = (v - vq) / 4;               // (4u + v) / 4 - u
            const T *col = sA + (vq * G::HA4 + u + vd) * G::PA;
            T av[G::LDA];
#pragma unroll
            for (int f = 0; f < G::LDA; ++f) av[f] = col[f];
            // this A column (p) pairs with output column w when e = p - q = vv - KUB ... :
            //   p = q0' + 4u + v, q = q0' + 4u + (HP-HQ) + w  =>  e = v - (HP-HQ) - w = vv - KUB - w
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int e = vv - KUB - w;
                if (e >= -KUB && e <= KLB) {
                    const T bc = bco[w][e + KUB];
                    // A[i,p], band row f  ->  AB band row r = U01 + i - q = U01 + (f - KUA) + e
#pragma unroll
                    for (int f = 0; f < G::LDA; ++f) acc[w][G::U01 - KUA + e + f] = fma(av[f], bc, acc[w][G::U01 - KUA + e + f]);
                }
            }
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int64_t q = q0 + 4 * u + w;
            const bool qok = q >= 0 && q < n;
#pragma unroll
            for (int r = 0; r < G::LD01; ++r) sAB[(w * G::HB4 + u) * G::P01 + r] = qok ? acc[w][r] : (T)0;
        }
    }
    __syncthreads();

    // ---- phase 2: D[:, j0 + 4t + w] = sum_e AB[:, j+e] * C[j+e, j] ---------------------------
    T dacc[4][G::LDD];
    if (tid < CH_NT) {
        const int t = tid;
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int r = 0; r < G::LDD; ++r) dacc[w][r] = (T)0;
        T cco[4][G::LDC];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int64_t j = j0 + 4 * t + w;
#pragma unroll
            for (int ep = 0; ep < G::LDC; ++ep) {
                const int64_t q = j + ep - KUC;
                cco[w][ep] = (j < n && q >= 0 && q < n) ? sC[(w * G::HC4 + t) * G::PC + ep] : (T)0;
            }
        }
        constexpr int V0 = G::HQ - KUC;
        constexpr int NV = 4 + KLC + KUC;
#pragma unroll
        for (int vv = 0; vv < NV; ++vv) {
            const int v = V0 + vv;
            const int vq = (v % 4 + 4) % 4;
            const int vd = (v - vq) / 4;
            const T *col = sAB + (vq * G::HB4 + t + vd) * G::P01;
            T av[G::LD01];
#pragma unroll
            for (int f = 0; f < G::LD01; ++f) av[f] = col[f];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int e = vv - KUC - w;
                if (e >= -KUC && e <= KLC) {
                    const T cc = cco[w][e + KUC];
                    // AB[i,q] band row f -> D band row r = U3 + i - j = U3 + (f - U01) + e
#pragma unroll
                    for (int f = 0; f < G::LD01; ++f) dacc[w][G::U3 - G::U01 + e + f] = fma(av[f], cc, dacc[w][G::U3 - G::U01 + e + f]);
                }
            }
        }
    }
    __syncthreads();  // everyone is done reading sC / sAB before D staging overwrites the A/B/C region
    if (tid < CH_NT) {
        const int t = tid;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int64_t j = j0 + 4 * t + w;
#pragma unroll
            for (int r = 0; r < G::LDD; ++r) {
                const int64_t i = j + r - G::U3;
                sD[(w * G::HC4 + t) * G::PD + r] = (i >= 0 && i < n) ? dacc[w][r] : (T)0;
            }
        }
    }
    __syncthreads();
    // coalesced copy-out of the contiguous D tile (128-bit stores when aligned)
    const int64_t ncol = j0 + CH_TCOL <= n ? CH_TCOL : n - j0;
    T *Dt = Dg + j0 * G::LDD;
    const int totD = (int)ncol * G::LDD;
    constexpr int EPVD = 16 / (int)sizeof(T);
    using VD = typename vec16<T>::type;
    if (totD % EPVD == 0 && lbm_aligned16_dev(Dt)) {
        for (int vi = tid; vi < totD / EPVD; vi += nthr) {
            VD v;
            T *pv = reinterpret_cast<T *>(&v);
#pragma unroll
            for (int w = 0; w < EPVD; ++w) {
                const int e = vi * EPVD + w;
                const int lc = e / G::LDD, r = e - lc * G::LDD;
                pv[w] = sD[((lc & 3) * G::HC4 + (lc >> 2)) * G::PD + r];
            }
            st_stream(reinterpret_cast<VD *>(Dt) + vi, v);
        }
    } else {
        for (int e = tid; e < totD; e += nthr) {
            const int lc = e / G::LDD, r = e - lc * G::LDD;
            Dt[e] = sD[((lc & 3) * G::HC4 + (lc >> 2)) * G::PD + r];
        }
    }
}

// ------------------------------------------------------------------------------------------
// "transposed" fused chain (variant for compile-time bands, Float32; chain_variant = 2/3): the shared operands A and
// A*B live in shared memory as [band row][column] so the 12 columns a thread needs of one band
// row are 3 LDS.128 (consecutive threads -> consecutive 16 bytes: conflict-free, 5x fewer loads);
// B's and C's coefficients for the thread's own 4 columns go global -> registers directly
// (LDB x LDG.128, no staging), and D leaves the registers as 25 x STG.128 (a thread's 4 columns
// are 4*LDD contiguous values).  Two __syncthreads per 512-column tile.
// ------------------------------------------------------------------------------------------
template <typename T, int KLA, int KUA, int KLB, int KUB, int KLC, int KUC, int MINB>
__global__ void __launch_bounds__(CH_NT + 32, MINB) gbmm_chain3_t_kernel(int64_t n, const T *__restrict__ A, int64_t sA_,
                                                                     const T *__restrict__ B, int64_t sB_,
                                                                     const T *__restrict__ C, int64_t sC_,
                                                                     T *__restrict__ D, int64_t sD_) {
    using G = ChainGeom<T, KLA, KUA, KLB, KUB, KLC, KUC>;
    static_assert(sizeof(T) == 4, "the transposed chain kernel is the Float32 path");
    constexpr int PAT = G::NP + 4;   // pitch of sA  [LDA ][NP]  (multiple of 4: LDS.128 aligned)
    constexpr int PQT = G::NQ;       // pitch of sAB [LD01][NQ]
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    T *sAB = sA + G::LDA * PAT;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int64_t b = blockIdx.y;
    const int64_t j0 = (int64_t)blockIdx.x * CH_TCOL;
    const int64_t q0 = j0 - G::HQ, p0 = j0 - G::HP;
    const T *Ag = A + b * sA_;
    const T *Bg = B + b * sB_;
    const T *Cg = C + b * sC_;
    T *Dg = D + b * sD_;
    // tiles that can touch the matrix boundary take the index-masked coefficient / store paths
    const bool edge = (p0 - G::L3 - G::U3 < 0) || (j0 + CH_TCOL + G::HP + G::L3 + G::U3 > n);

    // B and C tiles are consumed two phases later: start them towards L2 now so the coefficient
    // loads of phases 1 and 2 do not each pay a full DRAM latency
    {
        const int64_t qa = q0 < 0 ? 0 : q0, qb_ = q0 + G::NQ < n ? q0 + G::NQ : n;
        const char *bp = reinterpret_cast<const char *>(Bg + qa * G::LDB);
        const int64_t bbytes = (qb_ - qa) * G::LDB * (int64_t)sizeof(T);
        for (int64_t o = (int64_t)tid * 128; o < bbytes; o += (int64_t)nthr * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(bp + o));
        const int64_t je = j0 + CH_TCOL < n ? j0 + CH_TCOL : n;
        const char *cp = reinterpret_cast<const char *>(Cg + j0 * G::LDC);
        const int64_t cbytes = (je - j0) * G::LDC * (int64_t)sizeof(T);
        for (int64_t o = (int64_t)tid * 128; o < cbytes; o += (int64_t)nthr * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + o));
    }

    // ---- stage A (global, column-major) -> sA[f][col]; zero outside [0,n) -------------------
    {
        constexpr int total4 = G::NP * G::LDA / 4;
        const int64_t e_lo = (p0 < 0 ? -p0 : 0) * (int64_t)G::LDA;
        const int64_t c_end = p0 + G::NP < n ? p0 + G::NP : n;
        const int64_t e_hi = (c_end - p0) * (int64_t)G::LDA;
        const T *src = Ag + p0 * G::LDA;
        constexpr int VPT = (total4 + CH_NT + 31) / (CH_NT + 32);
        float4 v[VPT];
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int vi = k * nthr + tid;
            const int64_t e = (int64_t)vi * 4;
            v[k] = (vi < total4 && e >= e_lo && e + 4 <= e_hi) ? ld_stream(reinterpret_cast<const float4 *>(src + e)) : float4{0.f, 0.f, 0.f, 0.f};
        }
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int vi = k * nthr + tid;
            if (vi < total4) {
                const float *pv = reinterpret_cast<const float *>(&v[k]);
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int e = vi * 4 + w;
                    const int lc = e / G::LDA, f = e - lc * G::LDA;
                    sA[f * PAT + lc] = pv[w];
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 1: AB[:, q0 + 4u + w] for u = tid < NQ/4 --------------------------------------
    if (tid < G::NQ / 4) {
        const int u = tid;
        const int64_t qb = q0 + 4 * u;             // first of the thread's 4 columns (multiple of 4)
        const bool qok = qb >= 0 && qb < n;         // n % 4 == 0: all four columns valid or none
        T acc[4][G::LD01];
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int r = 0; r < G::LD01; ++r) acc[w][r] = (T)0;
        T bco[4 * G::LDB];                          // B band entries of the 4 columns, [w*LDB + ep]
        if (qok) {
#pragma unroll
            for (int k = 0; k < G::LDB; ++k) {
                const float4 t4 = ld_stream(reinterpret_cast<const float4 *>(Bg + qb * G::LDB) + k);
                bco[4 * k] = t4.x; bco[4 * k + 1] = t4.y; bco[4 * k + 2] = t4.z; bco[4 * k + 3] = t4.w;
            }
            if (edge) {
#pragma unroll
                for (int w = 0; w < 4; ++w)
#pragma unroll
                    for (int ep = 0; ep < G::LDB; ++ep) {
                        const int64_t pp = qb + w + ep - KUB;
                        if (pp < 0 || pp >= n) bco[w * G::LDB + ep] = (T)0;
                    }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4 * G::LDB; ++k) bco[k] = (T)0;
        }
        // A columns needed: local index 4u + v, v in [V0, V0 + 4 + KLB + KUB); loaded as aligned float4s
        constexpr int V0 = (G::HP - G::HQ) - KUB;
        constexpr int V0F = (V0 >= 0 ? V0 / 4 : -((-V0 + 3) / 4)) * 4;
        constexpr int NV4 = (V0 + 4 + KLB + KUB - V0F + 3) / 4;
#pragma unroll
        for (int f = 0; f < G::LDA; ++f) {
            T av[4 * NV4];
#pragma unroll
            for (int k = 0; k < NV4; ++k) {
                const float4 t4 = *reinterpret_cast<const float4 *>(sA + f * PAT + 4 * u + V0F + 4 * k);
                av[4 * k] = t4.x; av[4 * k + 1] = t4.y; av[4 * k + 2] = t4.z; av[4 * k + 3] = t4.w;
            }
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int e = -KUB; e <= KLB; ++e) {
                    // A column p = q + e  <->  v = (HP-HQ) + w + e;  AB band row r = U01 - KUA + e + f
                    acc[w][G::U01 - KUA + e + f] = fma(av[(G::HP - G::HQ) + w + e - V0F], bco[w * G::LDB + e + KUB], acc[w][G::U01 - KUA + e + f]);
                }
        }
#pragma unroll
        for (int r = 0; r < G::LD01; ++r) {
            float4 o;
            o.x = qok ? acc[0][r] : 0.f; o.y = qok ? acc[1][r] : 0.f; o.z = qok ? acc[2][r] : 0.f; o.w = qok ? acc[3][r] : 0.f;
            *reinterpret_cast<float4 *>(sAB + r * PQT + 4 * u) = o;
        }
    }
    __syncthreads();

    // ---- phase 2: D[:, j0 + 4t + w] = sum_e AB[:, j+e] * C[j+e, j] ---------------------------
    if (tid < CH_NT) {
        const int t = tid;
        const int64_t jb = j0 + 4 * t;
        if (jb >= n) return;                        // whole thread beyond the matrix (n % 4 == 0)
        T dacc[4][G::LDD];
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int r = 0; r < G::LDD; ++r) dacc[w][r] = (T)0;
        T cco[4 * G::LDC];
#pragma unroll
        for (int k = 0; k < G::LDC; ++k) {
            const float4 t4 = ld_stream(reinterpret_cast<const float4 *>(Cg + jb * G::LDC) + k);
            cco[4 * k] = t4.x; cco[4 * k + 1] = t4.y; cco[4 * k + 2] = t4.z; cco[4 * k + 3] = t4.w;
        }
        if (edge) {
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int ep = 0; ep < G::LDC; ++ep) {
                    const int64_t qq = jb + w + ep - KUC;
                    if (qq < 0 || qq >= n) cco[w * G::LDC + ep] = (T)0;
                }
        }
        constexpr int V0 = G::HQ - KUC;
        constexpr int V0F = (V0 >= 0 ? V0 / 4 : -((-V0 + 3) / 4)) * 4;
        constexpr int NV4 = (V0 + 4 + KLC + KUC - V0F + 3) / 4;
#pragma unroll
        for (int f = 0; f < G::LD01; ++f) {
            T av[4 * NV4];
#pragma unroll
            for (int k = 0; k < NV4; ++k) {
                const float4 t4 = *reinterpret_cast<const float4 *>(sAB + f * PQT + 4 * t + V0F + 4 * k);
                av[4 * k] = t4.x; av[4 * k + 1] = t4.y; av[4 * k + 2] = t4.z; av[4 * k + 3] = t4.w;
            }
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int e = -KUC; e <= KLC; ++e)
                    dacc[w][G::U3 - G::U01 + e + f] = fma(av[G::HQ + w + e - V0F], cco[w * G::LDC + e + KUC], dacc[w][G::U3 - G::U01 + e + f]);
        }
        // D: the thread's 4 columns are 4*LDD contiguous values -> LDD x STG.128
        T *Dt = Dg + jb * G::LDD;
#pragma unroll
        for (int k = 0; k < G::LDD; ++k) {
            float o[4];
#pragma unroll
            for (int z = 0; z < 4; ++z) {
                constexpr int dummy = 0;
                (void)dummy;
                const int idx = 4 * k + z;
                const int w = idx / G::LDD, r = idx - w * G::LDD;
                T val = dacc[w][r];
                if (edge) {
                    const int64_t i = jb + w + r - G::U3;
                    if (i < 0 || i >= n) val = (T)0;
                }
                o[z] = val;
            }
            st_stream(reinterpret_cast<float4 *>(Dt) + k, float4{o[0], o[1], o[2], o[3]});
        }
    }
}

template <typename T, int KLA, int KUA, int KLB, int KUB, int KLC, int KUC, int MINB>
int chain3_t_launch2(lbm_handle_t h, int64_t batch, int64_t n, const T *A, int64_t sA, const T *B, int64_t sB, const T *C,
                     int64_t sC, T *D, int64_t sD) {
    using G = ChainGeom<T, KLA, KUA, KLB, KUB, KLC, KUC>;
    constexpr size_t smem = (size_t)(G::LDA * (G::NP + 4) + G::LD01 * G::NQ) * sizeof(T);
    static std::atomic<int> configured_dev_mask{0};
    auto kern = gbmm_chain3_t_kernel<T, KLA, KUA, KLB, KUB, KLC, KUC, MINB>;
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_dev_mask |= (1 << h->device);
    }
    dim3 grid((unsigned)((n + CH_TCOL - 1) / CH_TCOL), (unsigned)batch);
    kern<<<grid, CH_NT + 32, smem, h->stream>>>(n, A, sA, B, sB, C, sC, D, sD);
    LBM_LAUNCH_CHECK(h, "gbmm_chain3_t_kernel");
    return 0;
}
template <typename T, int KLA, int KUA, int KLB, int KUB, int KLC, int KUC>
int chain3_t_launch(lbm_handle_t h, int64_t batch, int64_t n, const T *A, int64_t sA, const T *B, int64_t sB, const T *C,
                    int64_t sC, T *D, int64_t sD) {
    // chain_variant 2: 204 registers (2 CTAs/SM); 3: 128 registers (3 CTAs/SM)
    if (h->tune.chain_variant == 2)
        return chain3_t_launch2<T, KLA, KUA, KLB, KUB, KLC, KUC, 2>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
    return chain3_t_launch2<T, KLA, KUA, KLB, KUB, KLC, KUC, 3>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
}

template <typename T, int KLA, int KUA, int KLB, int KUB, int KLC, int KUC>
int chain3_reg_launch(lbm_handle_t h, int64_t batch, int64_t n, const T *A, int64_t sA, const T *B, int64_t sB,
                      const T *C, int64_t sC, T *D, int64_t sD) {
    using G = ChainGeom<T, KLA, KUA, KLB, KUB, KLC, KUC>;
    constexpr size_t smem = (size_t)G::SMEM_ELEMS * sizeof(T);
    static std::atomic<int> configured_dev_mask{0};
    auto kern = gbmm_chain3_reg_kernel<T, KLA, KUA, KLB, KUB, KLC, KUC>;
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_dev_mask |= (1 << h->device);
    }
    dim3 grid((unsigned)((n + CH_TCOL - 1) / CH_TCOL), (unsigned)batch);
    kern<<<grid, CH_NT + 32, smem, h->stream>>>(n, A, sA, B, sB, C, sC, D, sD);
    LBM_LAUNCH_CHECK(h, "gbmm_chain3_reg_kernel");
    return 0;
}

template <typename T>
int chain3_impl(lbm_handle_t h, int64_t batch, int64_t n, const int64_t *kl, const int64_t *ku, const T *A, int64_t sA,
                const T *B, int64_t sB, const T *C, int64_t sC, T *D, int64_t sD) {
    if (!h) return -1;
    if (batch < 0) return lbm_fail_arg(h, 2, "batch < 0");
    if (n < 0) return lbm_fail_arg(h, 3, "n < 0");
    if (!kl) return lbm_fail_arg(h, 4, "kl == NULL");
    if (!ku) return lbm_fail_arg(h, 5, "ku == NULL");
    for (int f = 0; f < 3; ++f) {
        if (kl[f] < 0) return lbm_fail_arg(h, 4, "kl[i] < 0");
        if (ku[f] < 0) return lbm_fail_arg(h, 5, "ku[i] < 0");
    }
    if (batch == 0 || n == 0) return 0;
    if (!A) return lbm_fail_arg(h, 6, "A == NULL");
    if (!B) return lbm_fail_arg(h, 8, "B == NULL");
    if (!C) return lbm_fail_arg(h, 10, "C == NULL");
    if (!D) return lbm_fail_arg(h, 12, "D == NULL");
    if (batch > 65535) return lbm_fail_arg(h, 2, "batch > 65535 (split the call)");
    lbm_device_guard g(h->device);
    if (!h->tune.chain_force_generic) {
        auto all = [&](int64_t l, int64_t u) {
            return kl[0] == l && kl[1] == l && kl[2] == l && ku[0] == u && ku[1] == u && ku[2] == u;
        };
        if constexpr (sizeof(T) == 4) {
            // chain_variant: 0 = auto / 4 = warp-autonomous TMA kernel (gbmm_chain_warp.cu) when the shape is covered,
            // 1 = split-column register kernel, 2/3 = transposed-smem kernel (204 / 128 registers)
            if (h->tune.chain_variant == 0 || h->tune.chain_variant == 4 || h->tune.chain_variant == 5) {
                bool handled = false;
                int rc = lbm_chain3_warp_f32(h, batch, n, kl, ku, A, sA, B, sB, C, sC, D, sD, &handled);
                if (handled) return rc;
            }
            const bool t_ok = (h->tune.chain_variant == 2 || h->tune.chain_variant == 3) && n % 4 == 0 && sA % 4 == 0 && sB % 4 == 0 && sC % 4 == 0 &&
                              sD % 4 == 0 && lbm_aligned16(A) && lbm_aligned16(B) && lbm_aligned16(C) && lbm_aligned16(D);
            if (t_ok && all(4, 4)) return chain3_t_launch<T, 4, 4, 4, 4, 4, 4>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
            if (t_ok && all(1, 1)) return chain3_t_launch<T, 1, 1, 1, 1, 1, 1>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
            if (t_ok && all(2, 2)) return chain3_t_launch<T, 2, 2, 2, 2, 2, 2>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
        }
        if (all(4, 4)) return chain3_reg_launch<T, 4, 4, 4, 4, 4, 4>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
        if (all(1, 1)) return chain3_reg_launch<T, 1, 1, 1, 1, 1, 1>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
        if (all(2, 2)) return chain3_reg_launch<T, 2, 2, 2, 2, 2, 2>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
    }
    ChainArgs<T> a;
    a.batch = batch; a.n = n;
    a.kla = (int)kl[0]; a.kua = (int)ku[0]; a.klb = (int)kl[1]; a.kub = (int)ku[1]; a.klc = (int)kl[2]; a.kuc = (int)ku[2];
    a.A = A; a.B = B; a.C = C; a.D = D; a.sA = sA; a.sB = sB; a.sC = sC; a.sD = sD;
    const int64_t lda = kl[0] + ku[0] + 1, ldb = kl[1] + ku[1] + 1, ldc = kl[2] + ku[2] + 1;
    const int64_t ld01 = kl[0] + kl[1] + ku[0] + ku[1] + 1;
    int64_t tile = h->tune.chain_tile > 0 ? h->tune.chain_tile : 128;
    auto layout = [&](int64_t tl, int64_t &oB, int64_t &oC, int64_t &oAB) {
        const int64_t qc = tl + kl[2] + ku[2];
        const int64_t pc = qc + kl[1] + ku[1];
        oB = pc * lda;
        oC = oB + qc * ldb;
        oAB = oC + tl * ldc;
        return (oAB + qc * ld01) * (int64_t)sizeof(T);
    };
    int64_t oB, oC, oAB;
    const int64_t smem_cap = 160 * 1024;
    while (tile > 8 && layout(tile, oB, oC, oAB) > smem_cap) tile /= 2;
    const int64_t smem = layout(tile, oB, oC, oAB);
    if (smem > smem_cap) return lbm_fail_arg(h, 4, "bands too wide for the fused chain kernel (use lbm_?gbmm twice)");
    a.tile = (int)tile; a.offB = (int)oB; a.offC = (int)oC; a.offAB = (int)oAB;
    static std::atomic<int> configured_dev_mask{0};
    auto kern = gbmm_chain3_kernel<T>;
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        configured_dev_mask |= (1 << h->device);
    }
    dim3 grid((unsigned)((n + tile - 1) / tile), (unsigned)batch);
    kern<<<grid, 256, (size_t)smem, h->stream>>>(a);
    LBM_LAUNCH_CHECK(h, "gbmm_chain3_kernel");
    return 0;
}

// element-wise widen / round (the Float32 wide-band route of lbm_sgbmm); grid-stride, 4 elements per thread per step
__global__ void __launch_bounds__(256) cvt_f32_f64_kernel(const float *__restrict__ src, double *__restrict__ dst, int64_t cnt) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += stride) dst[i] = (double)src[i];
}
__global__ void __launch_bounds__(256) cvt_f64_f32_kernel(const double *__restrict__ src, float *__restrict__ dst, int64_t cnt) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += stride) dst[i] = (float)src[i];
}

}  // namespace

extern "C" {
int lbm_dgbmm(lbm_handle_t h, int64_t m, int64_t n, int64_t k, double alpha, int64_t kla, int64_t kua, const double *A,
              int64_t lda, int64_t klb, int64_t kub, const double *B, int64_t ldb, double beta, int64_t klc,
              int64_t kuc, double *C, int64_t ldc) {
    if (h && h->tune.gbmm_force != 1) {
        bool handled = false;
        int rc = lbm_gbmm_dmma_f64(h, m, n, k, alpha, kla, kua, A, lda, klb, kub, B, ldb, beta, klc, kuc, C, ldc, &handled);
        if (handled) return rc;
    }
    return gbmm_impl<double>(h, m, n, k, alpha, kla, kua, A, lda, klb, kub, B, ldb, beta, klc, kuc, C, ldc);
}
int lbm_sgbmm(lbm_handle_t h, int64_t m, int64_t n, int64_t k, float alpha, int64_t kla, int64_t kua, const float *A,
              int64_t lda, int64_t klb, int64_t kub, const float *B, int64_t ldb, float beta, int64_t klc, int64_t kuc,
              float *C, int64_t ldc) {
    // Wide bands in Float32: there is no FP32 tensor-core mode with a 24-bit mantissa (TF32 has 10 bits, outside the parity
    // bound), and the diagonal-wise kernel runs at ~2 TFLOP/s.  Route through the FP64 DMMA tile kernel instead: widen A, B
    // (and C when beta != 0) into the handle's scratch, run lbm_gbmm_dmma_f64, round the band back.  Two extra streaming
    // passes (~12 bytes per stored element) buy a ~10x faster contraction, and the result is the correctly rounded f64 one.
    // Needs (lda*k + ldb*n + ldc*n) * 8 bytes of scratch: skipped (diagonal-wise kernel) when that is not available.
    if (h && h->tune.gbmm_force != 1 && m > 0 && n > 0 && k > 0 && A && B && C && kla >= 0 && kua >= 0 && klb >= 0 && kub >= 0 &&
        klc >= 0 && kuc >= 0 && lda >= kla + kua + 1 && ldb >= klb + kub + 1 && ldc == klc + kuc + 1 &&
        ((kla + kua + 1 >= 33 && klb + kub + 1 >= 33) || h->tune.gbmm_force == 2)) {
        const int64_t need_l = kla + klb < m - 1 ? kla + klb : m - 1, need_u = kua + kub < n - 1 ? kua + kub : n - 1;
        auto pad = [](int64_t e) { return (e + 1) / 2 * 2; };   // keep every f64 array 16-byte aligned
        const int64_t ea = pad(lda * k), eb = pad(ldb * n), ec = pad(ldc * n);
        const size_t bytes = (size_t)(ea + eb + ec) * sizeof(double);
        lbm_device_guard g(h->device);
        size_t free_b = 0, total_b = 0;
        const bool fits = bytes <= h->scratch_bytes || (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && bytes + (1u << 28) < free_b + h->scratch_bytes);
        if (klc >= need_l && kuc >= need_u && fits) {
            void *sc = nullptr;
            int rc = lbm_scratch(h, bytes, &sc);
            if (rc) return rc;
            double *A64 = (double *)sc, *B64 = A64 + ea, *C64 = B64 + eb;
            auto widen = [&](const float *src, double *dst, int64_t cnt) {
                int64_t blocks = (cnt + 1023) / 1024;
                if (blocks > (int64_t)h->sm_count * 16) blocks = (int64_t)h->sm_count * 16;
                cvt_f32_f64_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(src, dst, cnt);
            };
            widen(A, A64, lda * k);
            LBM_LAUNCH_CHECK(h, "cvt_f32_f64_kernel");
            widen(B, B64, ldb * n);
            LBM_LAUNCH_CHECK(h, "cvt_f32_f64_kernel");
            if (beta != 0.0f) {
                widen(C, C64, ldc * n);
                LBM_LAUNCH_CHECK(h, "cvt_f32_f64_kernel");
            } else {
                LBM_CUDA(h, cudaMemsetAsync(C64, 0, (size_t)(ldc * n) * sizeof(double), h->stream));
            }
            bool handled = false;
            rc = lbm_gbmm_dmma_f64(h, m, n, k, (double)alpha, kla, kua, A64, lda, klb, kub, B64, ldb, (double)beta, klc, kuc, C64, ldc, &handled);
            if (handled) {
                if (rc) return rc;
                int64_t blocks = (ldc * n + 1023) / 1024;
                if (blocks > (int64_t)h->sm_count * 16) blocks = (int64_t)h->sm_count * 16;
                cvt_f64_f32_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(C64, C, ldc * n);
                LBM_LAUNCH_CHECK(h, "cvt_f64_f32_kernel");
                return 0;
            }
            // not taken after all (the tensor path declined the shape): C is untouched, fall through
        }
    }
    return gbmm_impl<float>(h, m, n, k, alpha, kla, kua, A, lda, klb, kub, B, ldb, beta, klc, kuc, C, ldc);
}
int lbm_sgbmm_chain3_batched(lbm_handle_t h, int64_t batch, int64_t n, const int64_t *kl, const int64_t *ku,
                             const float *A, int64_t strideA, const float *B, int64_t strideB, const float *C,
                             int64_t strideC, float *D, int64_t strideD) {
    return chain3_impl<float>(h, batch, n, kl, ku, A, strideA, B, strideB, C, strideC, D, strideD);
}
int lbm_dgbmm_chain3_batched(lbm_handle_t h, int64_t batch, int64_t n, const int64_t *kl, const int64_t *ku,
                             const double *A, int64_t strideA, const double *B, int64_t strideB, const double *C,
                             int64_t strideC, double *D, int64_t strideD) {
    return chain3_impl<double>(h, batch, n, kl, ku, A, strideA, B, strideB, C, strideC, D, strideD);
}
}
