// diagtrav.cu -- DiagTrav ordering on the device: the anti-diagonal traversal that turns a coefficient matrix / cubic
// tensor into the block vector KronTrav operators act on (/root/reference/src/blockkron.jl:14-24 type, :101-133 block
// indexing; 3-D order pinned by test/test_blockkron.jl:53-57), its inverse (InvDiagTrav, :226-256) and
// `zero_bottomright` (:27-64), which the DiagTrav constructor applies.  Round 1 did these with eager torch indexing
// (torch.where, permute().contiguous(), index_select); they are pure data movement, one coalesced side each.
//
// 2-D, X m x n column-major (k fastest), mu = max(m,n), q = min(m,n): block K = 1..mu holds the entries with
// k + j = K + 1 (1-based), j ascending; its size is min(K, q), so it starts at
//     off2(K) = (K-1)K/2                      for K-1 <= q,      q(q+1)/2 + (K-1-q) q   otherwise;
// entries with k + j > mu + 1 are the zeroed bottom-right corner.
// 3-D, X n x n x n: block K = 1..n holds k + j + l = K + 2, k descending then j descending; size K(K+1)/2,
//     off3(K) = (K-1)K(K+1)/6,    position inside = (K-k)(K-k+1)/2 + (K+1-k) - j.
#include "lbm_common.cuh"

namespace {

__device__ __forceinline__ int64_t off2(int64_t K, int64_t q) {
    const int64_t t = K - 1;
    return t <= q ? t * (t + 1) / 2 : q * (q + 1) / 2 + (t - q) * q;
}

// MODE 0: v = Vector(DiagTrav(X))   MODE 1: X = InvDiagTrav(v) with the bottom-right corner zeroed   MODE 2: zero_bottomright!(X)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) diagtrav2_kernel(int64_t m, int64_t n, const T *__restrict__ src, T *__restrict__ dst) {
    const int64_t mu = m > n ? m : n, q = m < n ? m : n;
    const int64_t total = m * n, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t j = e / m, k = e - j * m;     // 0-based
        const int64_t K = k + j + 1;                // 1-based block
        const bool keep = K <= mu;
        if (MODE == 2) {
            if (!keep) dst[e] = (T)0;
            continue;
        }
        const int64_t K1m = K - m > 0 ? K - m : 0;   // j_min (0-based) of block K: max(0, K - m)
        const int64_t pos = off2(K, q) + (j - K1m);
        if (MODE == 0) {
            if (keep) dst[pos] = src[e];
        } else {
            dst[e] = keep ? src[pos] : (T)0;
        }
    }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256) diagtrav3_kernel(int64_t n, const T *__restrict__ src, T *__restrict__ dst) {
    const int64_t total = n * n * n, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t l = e / (n * n), r = e - l * n * n, j = r / n, k = r - j * n;   // 0-based
        const int64_t K = k + j + l + 1;            // 1-based block: (k+1)+(j+1)+(l+1) = K + 2
        const bool keep = K <= n;
        if (MODE == 2) {
            if (!keep) dst[e] = (T)0;
            continue;
        }
        const int64_t k1 = k + 1, j1 = j + 1;
        const int64_t pos = (K - 1) * K * (K + 1) / 6 + (K - k1) * (K - k1 + 1) / 2 + (K + 1 - k1) - j1;
        if (MODE == 0) {
            if (keep) dst[pos] = src[e];
        } else {
            dst[e] = keep ? src[pos] : (T)0;
        }
    }
}

template <typename T, int MODE>
int diagtrav_impl(lbm_handle_t h, int ndim, int64_t m, int64_t n, const T *src, T *dst) {
    if (!h) return -1;
    if (ndim != 2 && ndim != 3) return lbm_fail_arg(h, 2, "ndim must be 2 or 3");
    if (m < 0) return lbm_fail_arg(h, 3, "m < 0");
    if (n < 0) return lbm_fail_arg(h, 4, "n < 0");
    if (ndim == 3 && m != n) return lbm_fail_arg(h, 3, "3-D DiagTrav needs a cubic array (m == n)");
    const int64_t total = ndim == 2 ? m * n : n * n * n;
    if (total == 0) return 0;
    if (MODE != 2 && !src) return lbm_fail_arg(h, 5, "source == NULL");
    if (!dst) return lbm_fail_arg(h, 6, "destination == NULL");
    lbm_device_guard g(h->device);
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)h->sm_count * 16) blocks = (int64_t)h->sm_count * 16;
    if (ndim == 2)
        diagtrav2_kernel<T, MODE><<<(unsigned)blocks, 256, 0, h->stream>>>(m, n, src, dst);
    else
        diagtrav3_kernel<T, MODE><<<(unsigned)blocks, 256, 0, h->stream>>>(n, src, dst);
    LBM_LAUNCH_CHECK(h, "diagtrav_kernel");
    return 0;
}

}  // namespace

extern "C" {
int lbm_ddiagtrav(lbm_handle_t h, int ndim, int64_t m, int64_t n, const double *X, double *v) { return diagtrav_impl<double, 0>(h, ndim, m, n, X, v); }
int lbm_sdiagtrav(lbm_handle_t h, int ndim, int64_t m, int64_t n, const float *X, float *v) { return diagtrav_impl<float, 0>(h, ndim, m, n, X, v); }
int lbm_dinvdiagtrav(lbm_handle_t h, int ndim, int64_t m, int64_t n, const double *v, double *X) { return diagtrav_impl<double, 1>(h, ndim, m, n, v, X); }
int lbm_sinvdiagtrav(lbm_handle_t h, int ndim, int64_t m, int64_t n, const float *v, float *X) { return diagtrav_impl<float, 1>(h, ndim, m, n, v, X); }
int lbm_dzero_bottomright(lbm_handle_t h, int ndim, int64_t m, int64_t n, double *X) { return diagtrav_impl<double, 2>(h, ndim, m, n, nullptr, X); }
int lbm_szero_bottomright(lbm_handle_t h, int ndim, int64_t m, int64_t n, float *X) { return diagtrav_impl<float, 2>(h, ndim, m, n, nullptr, X); }
// length of Vector(DiagTrav(X)): pure integers
int64_t lbm_diagtrav_length(int ndim, int64_t m, int64_t n) {
    if (ndim == 3) return n * (n + 1) * (n + 2) / 6;
    if (ndim != 2 || m < 0 || n < 0) return -1;
    const int64_t mu = m > n ? m : n, q = m < n ? m : n;
    return mu <= q ? mu * (mu + 1) / 2 : q * (q + 1) / 2 + (mu - q) * q;
}
}
