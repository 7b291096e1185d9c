// reduce.cu -- single-pass reductions over Tridiagonal / SymTridiagonal / Bidiagonal operands
// (SURVEY.md 8(f) rank 3): the only explicit stencil loops of the reference besides mul!.
//
//   dot(x, A, y)   = sum_j (du[j-1] x[j-1] + d[j] x[j] + dl[j] x[j+1]) * y[j]
//                    (/root/reference/src/tridiag.jl:665-682; SymTridiagonal passes dl == du == ev,
//                     Bidiagonal passes a NULL diagonal, src/bidiag.jl:396-435 for its _sum)
//   _sum(A, dims)  : dims = 1 -> res[j] = du[j-1] + d[j] + dl[j]   (column sums, src/tridiag.jl:606-611)
//                    dims = 2 -> res[i] = dl[i-1] + d[i] + du[i]   (row sums,    src/tridiag.jl:612-617)
//                    dims = 0 -> sum(d) + sum(dl) + sum(du)        (Colon,       src/tridiag.jl:594-595)
//
// HBM-bound streaming kernels: every operand is read exactly once with coalesced loads (the +-1 shifted reads
// hit L1), persistent grid = SMs x 8 CTAs.  Scalars are reduced deterministically: warp shuffle tree ->
// one partial per CTA in scratch -> a second 1-CTA launch sums the partials in a fixed order (no atomics,
// bit-reproducible from run to run for a fixed grid).
#include "lbm_common.cuh"

namespace {

constexpr int RED_NT = 256;

template <typename T>
__device__ __forceinline__ T block_sum(T v, T *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    T r = (T)0;
    if (warp == 0) {
        r = lane < RED_NT / 32 ? sh[lane] : (T)0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;  // valid in thread 0
}

// partial[b] = sum over this CTA's rows of (x' A)[j] * y[j]
template <typename T>
__global__ void __launch_bounds__(RED_NT) tridiag_dot_kernel(int64_t n, const T *__restrict__ dl, const T *__restrict__ d,
                                                             const T *__restrict__ du, const T *__restrict__ x,
                                                             const T *__restrict__ y, T *__restrict__ partial) {
    __shared__ T sh[RED_NT / 32];
    const int64_t stride = (int64_t)gridDim.x * RED_NT;
    T acc = (T)0;
    for (int64_t j = (int64_t)blockIdx.x * RED_NT + threadIdx.x; j < n; j += stride) {
        T t = d[j] * x[j];
        if (du && j > 0) t = fma(du[j - 1], x[j - 1], t);
        if (dl && j + 1 < n) t = fma(dl[j], x[j + 1], t);
        acc = fma(t, y[j], acc);
    }
    const T r = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// partial[b] = sum of this CTA's share of d, dl, du
template <typename T>
__global__ void __launch_bounds__(RED_NT) tridiag_total_kernel(int64_t n, const T *__restrict__ dl, const T *__restrict__ d,
                                                               const T *__restrict__ du, T wl, T wu, T *__restrict__ partial) {
    __shared__ T sh[RED_NT / 32];
    const int64_t stride = (int64_t)gridDim.x * RED_NT;
    T acc = (T)0;
    for (int64_t j = (int64_t)blockIdx.x * RED_NT + threadIdx.x; j < n; j += stride) {
        T t = d[j];
        if (dl && j + 1 < n) t = fma(wl, dl[j], t);
        if (du && j + 1 < n) t = fma(wu, du[j], t);
        acc += t;
    }
    const T r = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

template <typename T>
__global__ void __launch_bounds__(RED_NT) final_sum_kernel(int np, const T *__restrict__ partial, T *__restrict__ out) {
    __shared__ T sh[RED_NT / 32];
    T acc = (T)0;
    for (int i = threadIdx.x; i < np; i += RED_NT) acc += partial[i];
    const T r = block_sum(acc, sh);
    if (threadIdx.x == 0) *out = r;
}

// dims = 1: res[j] = du[j-1] + d[j] + dl[j];  dims = 2: res[i] = dl[i-1] + d[i] + du[i]
template <typename T>
__global__ void __launch_bounds__(RED_NT) tridiag_dimsum_kernel(int64_t n, const T *__restrict__ lo, const T *__restrict__ d,
                                                                const T *__restrict__ hi, T *__restrict__ res) {
    // `lo` is the diagonal read at index j-1, `hi` the one read at index j (the caller swaps them for dims = 2)
    const int64_t stride = (int64_t)gridDim.x * RED_NT;
    for (int64_t j = (int64_t)blockIdx.x * RED_NT + threadIdx.x; j < n; j += stride) {
        T t = d[j];
        if (lo && j > 0) t += lo[j - 1];
        if (hi && j + 1 < n) t += hi[j];
        res[j] = t;
    }
}

template <typename T>
int reduce_grid(lbm_handle_t h, int64_t n) {
    int64_t g = (n + RED_NT - 1) / RED_NT;
    const int64_t cap = (int64_t)h->sm_count * 8;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

template <typename T>
int tridiag_dot_impl(lbm_handle_t h, int64_t n, const T *dl, const T *d, const T *du, const T *x, const T *y, T *result) {
    if (!h) return -1;
    if (n < 0) return lbm_fail_arg(h, 2, "n < 0");
    if (!result) return lbm_fail_arg(h, 8, "result == NULL");
    lbm_device_guard g(h->device);
    if (n == 0) {
        LBM_CUDA(h, cudaMemsetAsync(result, 0, sizeof(T), h->stream));
        return 0;
    }
    if (!d) return lbm_fail_arg(h, 4, "d == NULL");
    if (!x) return lbm_fail_arg(h, 6, "x == NULL");
    if (!y) return lbm_fail_arg(h, 7, "y == NULL");
    const int grid = reduce_grid<T>(h, n);
    void *scr;
    int rc = lbm_scratch(h, (size_t)grid * sizeof(T), &scr);
    if (rc) return rc;
    tridiag_dot_kernel<T><<<grid, RED_NT, 0, h->stream>>>(n, dl, d, du, x, y, (T *)scr);
    LBM_LAUNCH_CHECK(h, "tridiag_dot_kernel");
    final_sum_kernel<T><<<1, RED_NT, 0, h->stream>>>(grid, (const T *)scr, result);
    LBM_LAUNCH_CHECK(h, "final_sum_kernel");
    return 0;
}

template <typename T>
int tridiag_sum_impl(lbm_handle_t h, int64_t n, const T *dl, const T *d, const T *du, int dims, T *out) {
    if (!h) return -1;
    if (n < 0) return lbm_fail_arg(h, 2, "n < 0");
    if (dims < 0 || dims > 2) return lbm_fail_arg(h, 6, "dims must be 0 (all), 1 or 2");
    if (n == 0 && dims != 0) return 0;                    // empty result
    if (!out) return lbm_fail_arg(h, 7, "out == NULL");
    lbm_device_guard g(h->device);
    if (n == 0) {
        LBM_CUDA(h, cudaMemsetAsync(out, 0, sizeof(T), h->stream));
        return 0;
    }
    if (!d) return lbm_fail_arg(h, 4, "d == NULL");
    const int grid = reduce_grid<T>(h, n);
    if (dims == 0) {
        void *scr;
        int rc = lbm_scratch(h, (size_t)grid * sizeof(T), &scr);
        if (rc) return rc;
        // SymTridiagonal (dl == du) reads ev once and weights it by 2 (src/tridiag.jl:595)
        const bool sym = dl && dl == du;
        tridiag_total_kernel<T><<<grid, RED_NT, 0, h->stream>>>(n, dl, d, sym ? nullptr : du, sym ? (T)2 : (T)1, (T)1, (T *)scr);
        LBM_LAUNCH_CHECK(h, "tridiag_total_kernel");
        final_sum_kernel<T><<<1, RED_NT, 0, h->stream>>>(grid, (const T *)scr, out);
        LBM_LAUNCH_CHECK(h, "final_sum_kernel");
        return 0;
    }
    if (dims == 1) tridiag_dimsum_kernel<T><<<grid, RED_NT, 0, h->stream>>>(n, du, d, dl, out);
    else tridiag_dimsum_kernel<T><<<grid, RED_NT, 0, h->stream>>>(n, dl, d, du, out);
    LBM_LAUNCH_CHECK(h, "tridiag_dimsum_kernel");
    return 0;
}

}  // namespace

extern "C" {
int lbm_dtridiag_dot(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du, const double *x,
                     const double *y, double *result) {
    return tridiag_dot_impl<double>(h, n, dl, d, du, x, y, result);
}
int lbm_stridiag_dot(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du, const float *x,
                     const float *y, float *result) {
    return tridiag_dot_impl<float>(h, n, dl, d, du, x, y, result);
}
int lbm_dtridiag_sum(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du, int dims, double *out) {
    return tridiag_sum_impl<double>(h, n, dl, d, du, dims, out);
}
int lbm_stridiag_sum(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du, int dims, float *out) {
    return tridiag_sum_impl<float>(h, n, dl, d, du, dims, out);
}
}
