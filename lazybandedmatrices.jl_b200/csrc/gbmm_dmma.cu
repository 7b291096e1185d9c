// gbmm_dmma.cu -- wide-band FP64 banded x banded product on DMMA tensor-core tiles.
//
//   C[i,j] = sum_p A[i,p] B[p,j]      A m x k (kla,kua),  B k x n (klb,kub),  C band (klc,kuc)
//
// For wide bands (config C4: l=u=128, 257^2 = 66k MACs per output column, AI ~ 16 flop/B) the
// work is a real dense contraction over the parallelogram where the bands overlap, so it goes
// to the FP64 tensor pipe (`mma.sync.aligned.m8n8k4.f64`, SASS DMMA.8x8x4 -- tcgen05 has no
// FP64 kind, so DMMA is the Blackwell tensor path for Float64).
//
// Layout trick: in diagonal-major storage A[i,p] sits at data[(kua+i-p) + lda*p] =
// data[(kua+i) + (lda-1)*p], i.e. the band IS a dense column-major matrix with leading
// dimension lda-1 and a shifted row origin; out-of-band positions alias neighbouring columns
// and are masked BY INDEX while staging.
//
// Kernel: a CTA owns a 64x64 tile of C (dense coordinates) that intersects C's band; it sweeps
// the tile's p-range in chunks of 16, staging zero-padded dense 64x16 (A) and 16x64 (B) tiles in
// shared memory through a 3-stage cp.async pipeline (8-byte copies with zero-fill do the band
// masking; pitches 72 / 20 doubles make both the staging writes and the DMMA fragment loads
// bank-conflict-free), and 4 warps each
// keep a 32x32 block of C (16 DMMA accumulators) in registers.  8x8x4 sub-products whose A or B
// block lies wholly outside its band are skipped (warp-uniform test), which removes most of
// the padding waste of the parallelogram.
#include "lbm_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, KC = 16, NT = 128;
constexpr int PA = 72;  // sA[p][i], i fastest
constexpr int PB = 20;  // sB[j][p], p fastest

struct DmmaArgs {
    int64_t m, n, k;
    int64_t kla, kua, lda, klb, kub, ldb, klc, kuc, ldc;
    double alpha, beta;
    const double *A;
    const double *B;
    double *C;
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ int64_t floordiv(int64_t a, int64_t b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// 8-byte cp.async with zero-fill: copies `src` when ok, writes 0.0 otherwise (nothing is read).
__device__ __forceinline__ void cp_async8_zfill(double *dst_smem, const double *src, bool ok) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
    const int sz = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int NSTAGE = 3;
constexpr int STAGE_DBL = KC * PA + TN * PB;  // doubles per stage

__global__ void __launch_bounds__(NT, 4) gbmm_dmma_kernel(const DmmaArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wi = (warp & 1) * 32, wj = (warp >> 1) * 32;  // warp's 32x32 block inside the tile

    const int64_t j0 = (int64_t)blockIdx.x * TN;
    int64_t I0 = floordiv(j0 - a.kuc, TM);
    if (I0 < 0) I0 = 0;
    const int64_t i0 = (I0 + blockIdx.y) * TM;
    if (i0 >= a.m || i0 > j0 + TN - 1 + a.klc) return;

    // p-range of the tile
    int64_t plo = i0 - a.kla > j0 - a.kub ? i0 - a.kla : j0 - a.kub;
    int64_t phi = i0 + TM - 1 + a.kua < j0 + TN - 1 + a.klb ? i0 + TM - 1 + a.kua : j0 + TN - 1 + a.klb;
    if (plo < 0) plo = 0;
    if (phi > a.k - 1) phi = a.k - 1;
    plo &= ~(int64_t)3;
    const int nch = phi >= plo ? (int)((phi - plo + KC) / KC) : 0;

    double acc[4][4][2];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;

    // ---- staging maps (all offsets relative to the tile: 32-bit) ---------------------------
    // A: ii = tid % 64, pp = tid / 64 + 2q (q < 8).  A[i,p] = Ad[(kua + i) + (lda-1)*p], valid for
    //    p in [i - kla, i + kua], p < k, i < m.
    // B: pp = tid % 16, jj = tid / 16 + 8q (q < 8).  B[p,j] = Bd[(kub + p) + (ldb-1)*j], valid for
    //    p in [j - kub, j + klb], p < k, j < n.
    const int a_ii = tid & 63, a_p0 = tid >> 6;
    const int b_pp = tid & 15, b_j0 = tid >> 4;
    const bool a_row_ok = (i0 + a_ii) < a.m;
    const int a_lo = (int)(i0 + a_ii - a.kla - plo);  // valid relative p range for this thread's row
    const int a_hi = (int)(i0 + a_ii + a.kua - plo);
    const int kmax = (int)(a.k - 1 - plo);             // relative p must be <= kmax
    const double *a_ptr = a.A + (a.kua + i0 + a_ii) + (a.lda - 1) * plo;  // + (lda-1)*prel
    const int64_t a_step = a.lda - 1;
    int b_lo[8], b_hi[8];
    bool b_col_ok[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int64_t j = j0 + b_j0 + 8 * q;
        b_col_ok[q] = j < a.n;
        b_lo[q] = (int)(j - a.kub - plo);
        b_hi[q] = (int)(j + a.klb - plo);
    }
    const double *b_ptr = a.B + (a.kub + plo + b_pp) + (a.ldb - 1) * (j0 + b_j0);  // + prel + (ldb-1)*8q
    const int64_t b_step = (a.ldb - 1) * 8;

    auto issue = [&](int ch) {
        double *sA = smem + (ch % NSTAGE) * STAGE_DBL;
        double *sB = sA + KC * PA;
        const int pr = ch * KC;  // relative p of the chunk start
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int p = pr + a_p0 + 2 * q;
            const bool ok = a_row_ok && p >= a_lo && p <= a_hi && p <= kmax;
            cp_async8_zfill(sA + (a_p0 + 2 * q) * PA + a_ii, ok ? a_ptr + a_step * p : a.A, ok);
        }
        const int pb = pr + b_pp;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const bool ok = b_col_ok[q] && pb >= b_lo[q] && pb <= b_hi[q] && pb <= kmax;
            cp_async8_zfill(sB + (b_j0 + 8 * q) * PB + b_pp, ok ? b_ptr + pr + b_step * q : a.B, ok);
        }
    };

    // prologue: NSTAGE-1 chunks in flight
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (s < nch) issue(s);
        cp_async_commit();
    }

    const int fr = lane >> 2, fc = lane & 3;  // fragment row / col of this lane
    // skip-test constants: block x of A covers rows ia..ia+7, block y of B covers cols jb..jb+7
    const int ia0 = (int)(i0 + wi - plo), jb0 = (int)(j0 + wj - plo);
    const int kla = (int)a.kla, kua = (int)a.kua, klb = (int)a.klb, kub = (int)a.kub;

    for (int ch = 0; ch < nch; ++ch) {
        cp_async_wait<NSTAGE - 2>();  // chunk ch has landed (this thread's copies)
        __syncthreads();              // ... everyone's; and everyone is done computing chunk ch-1
        if (ch + NSTAGE - 1 < nch) issue(ch + NSTAGE - 1);
        cp_async_commit();
        const double *sA = smem + (ch % NSTAGE) * STAGE_DBL;
        const double *sB = sA + KC * PA;
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            const int p = ch * KC + 4 * ks;  // relative
            unsigned ma = 0, mb = 0;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int ia = ia0 + 8 * x;
                if (ia - (p + 3) <= kla && (ia + 7) - p >= -kua) ma |= 1u << x;
                const int jb = jb0 + 8 * x;
                if ((p + 3) - jb >= -kub && p - (jb + 7) <= klb) mb |= 1u << x;
            }
            if (ma == 0 || mb == 0) continue;
            double fa[4], fb[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                fa[x] = sA[(4 * ks + fc) * PA + wi + 8 * x + fr];
                fb[x] = sB[(wj + 8 * x + fr) * PB + 4 * ks + fc];
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
                    if ((ma >> x & 1u) && (mb >> y & 1u)) dmma884(acc[x][y][0], acc[x][y][1], fa[x], fb[y]);
        }
    }
    cp_async_wait<0>();

    // epilogue: C fragment (row = lane/4, cols = 2*(lane%4) + {0,1}) -> band storage
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t i = i0 + wi + 8 * x + fr;
                const int64_t j = j0 + wj + 8 * y + 2 * fc + e;
                const int64_t f = a.kuc + i - j;
                if (i < a.m && j < a.n && f >= 0 && f <= a.klc + a.kuc) {
                    double *c = a.C + f + a.ldc * j;
                    const double v = a.alpha * acc[x][y][e];
                    *c = (a.beta == 0.0) ? v : fma(a.beta, *c, v);
                }
            }
}

// out-of-matrix slots of C's band (first kuc / last klc columns) are written 0 when beta == 0
__global__ void __launch_bounds__(256) gbmm_zero_corners_kernel(int64_t m, int64_t n, int64_t klc, int64_t kuc,
                                                               double *C, int64_t ldc) {
    const int64_t nb = klc + kuc + 1;
    const int64_t ncols_top = kuc < n ? kuc : n;  // columns j < kuc have slots with i < 0
    int64_t jbot = m - klc - 1;                   // columns j > m-klc-1 have slots with i >= m
    if (jbot < 0) jbot = 0;
    const int64_t ncols_bot = n > jbot ? n - jbot : 0;
    const int64_t total = (ncols_top + ncols_bot) * nb;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t cidx = e / nb, r = e - cidx * nb;
        const int64_t j = cidx < ncols_top ? cidx : jbot + (cidx - ncols_top);
        const int64_t i = j + r - kuc;
        if (i < 0 || i >= m) C[r + ldc * j] = 0.0;
    }
}

}  // namespace

int lbm_gbmm_dmma_f64(lbm_handle_t h, int64_t m, int64_t n, int64_t k, double alpha, int64_t kla, int64_t kua,
                      const double *A, int64_t lda, int64_t klb, int64_t kub, const double *B, int64_t ldb,
                      double beta, int64_t klc, int64_t kuc, double *C, int64_t ldc, bool *handled) {
    *handled = false;
    if (!h || m <= 0 || n <= 0 || k <= 0 || !A || !B || !C) return 0;  // generic path reports the error
    if (kla < 0 || kua < 0 || klb < 0 || kub < 0 || klc < 0 || kuc < 0) return 0;
    if (lda < kla + kua + 1 || ldb < klb + kub + 1 || ldc < klc + kuc + 1) return 0;
    const int64_t need_l = kla + klb < m - 1 ? kla + klb : m - 1;
    const int64_t need_u = kua + kub < n - 1 ? kua + kub : n - 1;
    if (klc < need_l || kuc < need_u) return 0;
    // tensor path pays off once both bands are wide (>= ~32 diagonals); forced with gbmm_force = 2
    const bool wide = (kla + kua + 1) >= 33 && (klb + kub + 1) >= 33;
    if (!(wide || h->tune.gbmm_force == 2)) return 0;
    *handled = true;
    lbm_device_guard g(h->device);
    DmmaArgs a;
    a.m = m; a.n = n; a.k = k; a.kla = kla; a.kua = kua; a.lda = lda; a.klb = klb; a.kub = kub; a.ldb = ldb;
    a.klc = klc; a.kuc = kuc; a.ldc = ldc; a.alpha = alpha; a.beta = beta; a.A = A; a.B = B; a.C = C;
    const int64_t gx = (n + TN - 1) / TN;
    int64_t gy = (klc + kuc + TN - 1) / TM + 2;
    const int64_t mt = (m + TM - 1) / TM;
    if (gy > mt) gy = mt;
    if (gy > 65535) return lbm_fail_arg(h, 15, "C band too wide for the DMMA tile grid");
    if (beta == 0.0) {
        gbmm_zero_corners_kernel<<<64, 256, 0, h->stream>>>(m, n, klc, kuc, C, ldc);
        LBM_LAUNCH_CHECK(h, "gbmm_zero_corners_kernel");
    }
    dim3 grid((unsigned)gx, (unsigned)gy);
    const size_t smem = (size_t)NSTAGE * STAGE_DBL * sizeof(double);
    static int configured_dev_mask = 0;
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_dev_mask |= (1 << h->device);
    }
    gbmm_dmma_kernel<<<grid, NT, smem, h->stream>>>(a);
    LBM_LAUNCH_CHECK(h, "gbmm_dmma_kernel");
    return 0;
}
