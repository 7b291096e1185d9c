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
// One fused kernel: a CTA owns a TR x TC tile of Y, stages the (TR+klb+kub) x (TC+kla+kua)
// halo'd tile of X in shared memory (zero outside the grid), stages the tile's coefficient
// rows of A (per column) and B (per row) with out-of-matrix slots masked to zero BY INDEX, then
// each thread walks its rows across the tile's columns.
#include "lbm_common.cuh"

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
    const int64_t c0 = (int64_t)blockIdx.y * a.tc;
    const int trn = (int)(r0 + a.tr < a.n2 ? a.tr : a.n2 - r0);
    const int tcn = (int)(c0 + a.tc < a.n1 ? a.tc : a.n1 - c0);
    const int hl = a.useB ? a.klb : 0, hu = a.useB ? a.kub : 0;   // row halo
    const int wl = a.useA ? a.kla : 0, wu = a.useA ? a.kua : 0;   // column halo
    const int rows_h = trn + hl + hu;
    const int cols_h = tcn + wl + wu;

    // X tile with halo; zero outside the grid
    for (int e = tid; e < rows_h * cols_h; e += 256) {
        const int cc = e / rows_h;
        const int rr = e - cc * rows_h;
        const int64_t r = r0 - hl + rr;
        const int64_t c = c0 - wl + cc;
        T v = 0;
        if (r >= 0 && r < a.n2 && c >= 0 && c < a.n1) v = a.x[r + a.n2 * c];
        sX[rr + a.pitch * cc] = v;
    }
    if (a.useA)
        for (int e = tid; e < tcn * nba; e += 256) {
            const int cc = e / nba;
            const int t = e - cc * nba;
            const int64_t c = c0 + cc;
            const int64_t q = c - a.kla + t;  // A[c,q] = Ad[(kua + c - q) + lda*q]
            sCA[e] = (q >= 0 && q < a.n1) ? a.A[(a.kua + c - q) + a.lda * q] : (T)0;
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
            a.y[o] = (a.beta == (T)0) ? a.alpha * acc : fma(a.beta, a.y[o], a.alpha * acc);
        }
    }
}

template <typename T>
int kron_launch(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua, const T *A, int64_t lda, int64_t klb,
                int64_t kub, const T *B, int64_t ldb, T alpha, const T *x, T beta, T *y, int useA, int useB) {
    KronArgs<T> a;
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
    const int64_t gx = (n2 + tr - 1) / tr, gy = (n1 + tc - 1) / tc;
    if (gy > 65535) return lbm_fail_arg(h, 2, "n1 too large for the tile grid");
    static int configured_dev_mask = 0;
    auto kern = kron_tile_kernel<T>;
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        configured_dev_mask |= (1 << h->device);
    }
    dim3 grid((unsigned)gx, (unsigned)gy);
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
    return kron_launch<T>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, x, beta, y, 1, 1);
}

}  // namespace

extern "C" {
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
    int rc = kron_check<double>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, x, y);
    if (rc || n1 == 0 || n2 == 0) return rc;
    lbm_device_guard g(h->device);
    void *z;
    if ((rc = lbm_scratch(h, (size_t)(n1 * n2) * sizeof(double), &z))) return rc;
    // Z = B*X, then Y = alpha*Z*A^T + beta*Y  (left to right, as src/blockkron.jl:440 evaluates B*X*A')
    rc = kron_launch<double>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, 1.0, x, 0.0, (double *)z, 0, 1);
    if (rc) return rc;
    return kron_launch<double>(h, n1, n2, kla, kua, A, lda, klb, kub, B, ldb, alpha, (const double *)z, beta, y, 1, 0);
}
}
