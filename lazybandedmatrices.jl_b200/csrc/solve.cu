// solve.cu -- X <- A \ B for the structured kinds: Tridiagonal / SymTridiagonal (dl == du) / Bidiagonal (one
// off-diagonal NULL: a triangular solve).  SURVEY.md 8(f) rank 4, exercised by the reference at
// /root/reference/test/test_tridiag.jl:345-363 (`A90\b90 ≈ inv(A90)*b90`) -- in the reference `\` falls through
// ArrayLayouts to LinearAlgebra's tridiagonal LU, a strictly sequential recurrence in n.
//
// HBM-bound like everything else on this path, but with a loop-carried dependence, so the algorithm is the partition
// method (Wang 1981) applied recursively:
//   pass 1  every THREAD owns R = 8 consecutive rows IN REGISTERS (a tile of 256 x 8 rows is staged through shared memory
//           with coalesced loads, chunk pitch R+1 -> conflict-free; 16 warps per SM hide the dependent-division latency,
//           one reciprocal per row).  Forward elimination inside the chunk expresses each row through u_{c-1} (= x of the
//           previous chunk's last row) and its own unknowns; a backward sweep expresses the chunk's FIRST row through
//           u_{c-1} and u_c.  Only 8 numbers per chunk leave the SM.
//   reduce  combining the last row of chunk c with the first row of chunk c+1 gives a scalar tridiagonal system of
//           size P = ceil(n/R) in the u_c -- solved by the same three kernels (levels shrink 8x).
//   pass 2  with u_{c-1}, u_c known every chunk is an independent (R-1)-row system: plain Thomas in registers,
//           coalesced store of x through shared memory.
// Traffic: 2 reads of (dl, d, du, b) + 1 write of x = 9n elements, + ~3.3 per row for the reduced levels, against the 5n of
// a single pass: the ceiling of this formulation is ~40 % of the copy roofline.  NO PIVOTING (as gtsv2_nopivot / Thomas):
// exact for diagonally dominant and symmetric positive definite operators -- the discretisation matrices this package
// exists for; a zero pivot is reported through *info (LAPACK style, > 0) instead of being hidden.
#include "lbm_common.cuh"

namespace {

constexpr int SV_R = 8;      // rows per thread chunk
constexpr int SV_NT = 256;   // threads per CTA
constexpr int SV_TILE = SV_R * SV_NT;
constexpr int SV_PITCH = SV_R + 1;

__device__ __forceinline__ double rcp_(double x) { return __drcp_rn(x); }
__device__ __forceinline__ float rcp_(float x) { return __frcp_rn(x); }

template <typename T>
struct SolveArgs {
    int64_t n;
    const T *dl, *d, *du;   // dl[i-1] = A[i,i-1], du[i] = A[i,i+1]; NULL = that diagonal is absent
    const T *b;             // right-hand side (one column)
    T *x;                   // pass 2: solution
    T *red;                 // pass 1: 8 arrays of P: fL dL cL gL fF dF rF gF
    const T *u;             // pass 2: u_c = x[last row of chunk c]
    int64_t P;              // chunks
    int *info;              // device flag: set to 1 on a zero pivot (may be NULL)
};

// stage rows [t0, t0 + SV_TILE) as (a, b, c, g); rows >= n are identity rows (decoupled, solution 0)
template <typename T>
__device__ __forceinline__ void stage_tile(const SolveArgs<T> &a, int64_t t0, T *sa, T *sb, T *sc, T *sg) {
#pragma unroll
    for (int q = 0; q < SV_R; ++q) {
        const int idx = threadIdx.x + q * SV_NT;
        const int64_t i = t0 + idx;
        const int s = (idx / SV_R) * SV_PITCH + (idx % SV_R);
        const bool in = i < a.n;
        sa[s] = (in && i > 0 && a.dl) ? a.dl[i - 1] : (T)0;
        sb[s] = in ? a.d[i] : (T)1;
        sc[s] = (in && i < a.n - 1 && a.du) ? a.du[i] : (T)0;
        sg[s] = in ? a.b[i] : (T)0;
    }
}

template <typename T>
__global__ void __launch_bounds__(SV_NT, 2) tridiag_solve_pass1_kernel(const SolveArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sa = reinterpret_cast<T *>(smem_raw);
    T *sb = sa + SV_NT * SV_PITCH, *sc = sb + SV_NT * SV_PITCH, *sg = sc + SV_NT * SV_PITCH;
    const int64_t ntiles = (a.n + SV_TILE - 1) / SV_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t t0 = tile * SV_TILE;
        __syncthreads();
        stage_tile<T>(a, t0, sa, sb, sc, sg);
        __syncthreads();
        const int64_t c = tile * SV_NT + threadIdx.x;   // global chunk
        if (c < a.P) {
            const int o = threadIdx.x * SV_PITCH;
            T f[SV_R], d[SV_R], cc[SV_R], g[SV_R], ri[SV_R];
#pragma unroll
            for (int k = 0; k < SV_R; ++k) { f[k] = sa[o + k]; d[k] = sb[o + k]; cc[k] = sc[o + k]; g[k] = sg[o + k]; }
            // forward: row k becomes  f_k u_{c-1} + d_k x_k + c_k x_{k+1} = g_k
            bool bad = d[0] == (T)0;
            ri[0] = rcp_(d[0]);
#pragma unroll
            for (int k = 1; k < SV_R; ++k) {
                const T m = f[k] * ri[k - 1];
                f[k] = -m * f[k - 1];
                d[k] = d[k] - m * cc[k - 1];
                g[k] = g[k] - m * g[k - 1];
                bad |= d[k] == (T)0;
                ri[k] = rcp_(d[k]);
            }
            // backward: the FIRST row becomes  fF u_{c-1} + dF x_first + rF u_c = gF
            T f2 = f[SV_R - 2], r = cc[SV_R - 2], g2 = g[SV_R - 2];
#pragma unroll
            for (int k = SV_R - 3; k >= 0; --k) {
                const T m = cc[k] * ri[k + 1];
                f2 = f[k] - m * f2;
                r = -m * r;
                g2 = g[k] - m * g2;
            }
            T *out = a.red + c;
            out[0] = f[SV_R - 1]; out[a.P] = d[SV_R - 1]; out[2 * a.P] = cc[SV_R - 1]; out[3 * a.P] = g[SV_R - 1];
            out[4 * a.P] = f2; out[5 * a.P] = d[0]; out[6 * a.P] = r; out[7 * a.P] = g2;
            if (bad && a.info) *a.info = 1;
        }
    }
}

// reduced system in the u_c: eliminate x_first of chunk c+1 from the last row of chunk c
//   fL_c u_{c-1} + dL_c u_c + cL_c x_first(c+1) = gL_c,   x_first(c+1) = (gF - fF u_c - rF u_{c+1}) / dF   [of chunk c+1]
template <typename T>
__global__ void __launch_bounds__(256) tridiag_solve_reduce_kernel(int64_t P, const T *red, T *ra, T *rb, T *rc, T *rg, int *info) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < P; c += stride) {
        const T fL = red[c], dL = red[P + c], cL = red[2 * P + c], gL = red[3 * P + c];
        T bb = dL, cc = (T)0, gg = gL;
        if (c + 1 < P) {
            const T fF = red[4 * P + c + 1], dF = red[5 * P + c + 1], rF = red[6 * P + c + 1], gF = red[7 * P + c + 1];
            if (dF == (T)0 && cL != (T)0 && info) *info = 1;
            const T q = cL == (T)0 ? (T)0 : cL / dF;
            bb = dL - q * fF;
            cc = -q * rF;
            gg = gL - q * gF;
        }
        ra[c] = fL;   // coupling to u_{c-1}: the sub-diagonal of the reduced system is ra[1..P)
        rb[c] = bb;
        rc[c] = cc;
        rg[c] = gg;
    }
}

template <typename T>
__global__ void __launch_bounds__(SV_NT, 2) tridiag_solve_pass2_kernel(const SolveArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sa = reinterpret_cast<T *>(smem_raw);
    T *sb = sa + SV_NT * SV_PITCH, *sc = sb + SV_NT * SV_PITCH, *sg = sc + SV_NT * SV_PITCH;
    const int64_t ntiles = (a.n + SV_TILE - 1) / SV_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t t0 = tile * SV_TILE;
        __syncthreads();
        stage_tile<T>(a, t0, sa, sb, sc, sg);
        __syncthreads();
        const int64_t c = tile * SV_NT + threadIdx.x;
        if (c < a.P) {
            const int o = threadIdx.x * SV_PITCH;
            T av[SV_R], bv[SV_R], cp[SV_R], gp[SV_R];
#pragma unroll
            for (int k = 0; k < SV_R; ++k) { av[k] = sa[o + k]; bv[k] = sb[o + k]; cp[k] = sc[o + k]; gp[k] = sg[o + k]; }
            const T uprev = c > 0 ? a.u[c - 1] : (T)0;
            const T uc = a.u[c];
            // rows 0 .. R-2 with x_{-1} = uprev and x_{R-1} = uc known: Thomas
            T r = rcp_(bv[0]);
            cp[0] = cp[0] * r;
            gp[0] = (gp[0] - av[0] * uprev) * r;
#pragma unroll
            for (int k = 1; k < SV_R - 1; ++k) {
                r = rcp_(bv[k] - av[k] * cp[k - 1]);
                cp[k] = cp[k] * r;
                gp[k] = (gp[k] - av[k] * gp[k - 1]) * r;
            }
            T xk = uc;
            sg[o + SV_R - 1] = xk;
#pragma unroll
            for (int k = SV_R - 2; k >= 0; --k) {
                xk = gp[k] - cp[k] * xk;
                sg[o + k] = xk;
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < SV_R; ++q) {
            const int idx = threadIdx.x + q * SV_NT;
            const int64_t i = t0 + idx;
            if (i < a.n) a.x[i] = sg[(idx / SV_R) * SV_PITCH + (idx % SV_R)];
        }
    }
}

// base of the recursion: one thread, plain Thomas (n <= a few hundred)
template <typename T>
__global__ void tridiag_solve_seq_kernel(int64_t n, const T *dl, const T *d, const T *du, const T *b, T *x, T *work, int *info) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // work: c' (n)
    T den = d[0];
    bool bad = den == (T)0;
    T cp = (du && n > 1) ? du[0] / den : (T)0, gp = b[0] / den;
    work[0] = cp;
    x[0] = gp;
    for (int64_t i = 1; i < n; ++i) {
        const T ai = dl ? dl[i - 1] : (T)0;
        den = d[i] - ai * cp;
        bad |= den == (T)0;
        cp = (du && i < n - 1) ? du[i] / den : (T)0;
        gp = (b[i] - ai * gp) / den;
        work[i] = cp;
        x[i] = gp;
    }
    for (int64_t i = n - 2; i >= 0; --i) x[i] = x[i] - work[i] * x[i + 1];
    if (bad && info) *info = 1;
}

constexpr int64_t SV_BASE = 32;

// scratch elements one right-hand side needs below a system of size n (all levels)
int64_t solve_scratch_elems(int64_t n) {
    int64_t tot = 0;
    while (n > SV_BASE) {
        const int64_t P = (n + SV_R - 1) / SV_R;
        tot += 13 * P + 64;   // red (8P) + reduced a,b,c,g (4P) + u (P)
        n = P;
    }
    return tot + SV_BASE + 64;
}

template <typename T>
int solve_level(lbm_handle_t h, int64_t n, const T *dl, const T *d, const T *du, const T *b, T *x, T *scratch, int *info) {
    if (n <= SV_BASE) {
        tridiag_solve_seq_kernel<T><<<1, 32, 0, h->stream>>>(n, dl, d, du, b, x, scratch, info);
        LBM_LAUNCH_CHECK(h, "tridiag_solve_seq_kernel");
        return 0;
    }
    const int64_t P = (n + SV_R - 1) / SV_R;
    T *red = scratch, *ra = red + 8 * P, *rb = ra + P, *rc = rb + P, *rg = rc + P, *u = rg + P;
    T *next = u + P + ((16 - (13 * P) % 16) % 16);   // keep the next level 16-element aligned
    SolveArgs<T> a;
    a.n = n; a.dl = dl; a.d = d; a.du = du; a.b = b; a.x = x; a.red = red; a.u = u; a.P = P; a.info = info;
    const size_t smem = (size_t)4 * SV_NT * SV_PITCH * sizeof(T);
    static std::atomic<int> configured_dev_mask{0};
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(tridiag_solve_pass1_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LBM_CUDA(h, cudaFuncSetAttribute(tridiag_solve_pass2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_dev_mask |= (1 << h->device);
    }
    const int64_t ntiles = (n + SV_TILE - 1) / SV_TILE;
    int64_t grid = (int64_t)h->sm_count * 2;
    if (grid > ntiles) grid = ntiles;
    tridiag_solve_pass1_kernel<T><<<(unsigned)grid, SV_NT, smem, h->stream>>>(a);
    LBM_LAUNCH_CHECK(h, "tridiag_solve_pass1_kernel");
    int64_t rblocks = (P + 255) / 256;
    if (rblocks > (int64_t)h->sm_count * 8) rblocks = (int64_t)h->sm_count * 8;
    tridiag_solve_reduce_kernel<T><<<(unsigned)rblocks, 256, 0, h->stream>>>(P, red, ra, rb, rc, rg, info);
    LBM_LAUNCH_CHECK(h, "tridiag_solve_reduce_kernel");
    int rc2 = solve_level<T>(h, P, ra + 1, rb, rc, rg, u, next, info);
    if (rc2) return rc2;
    tridiag_solve_pass2_kernel<T><<<(unsigned)grid, SV_NT, smem, h->stream>>>(a);
    LBM_LAUNCH_CHECK(h, "tridiag_solve_pass2_kernel");
    return 0;
}

template <typename T>
int tridiag_solve_impl(lbm_handle_t h, int64_t n, const T *dl, const T *d, const T *du, const T *B, int64_t ldb, int64_t nrhs,
                       T *X, int64_t ldx, int *info) {
    if (!h) return -1;
    if (n < 0) return lbm_fail_arg(h, 2, "n < 0");
    if (nrhs < 0) return lbm_fail_arg(h, 8, "nrhs < 0");
    if (n == 0 || nrhs == 0) return 0;
    if (!d) return lbm_fail_arg(h, 4, "d == NULL");
    if (!B) return lbm_fail_arg(h, 6, "B == NULL");
    if (ldb < n && nrhs > 1) return lbm_fail_arg(h, 7, "ldb < n");
    if (!X) return lbm_fail_arg(h, 9, "X == NULL");
    if (ldx < n && nrhs > 1) return lbm_fail_arg(h, 10, "ldx < n");
    lbm_device_guard g(h->device);
    void *scr;
    int rc = lbm_scratch(h, (size_t)(solve_scratch_elems(n) + 64) * sizeof(T), &scr);
    if (rc) return rc;
    if (info) LBM_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int), h->stream));
    for (int64_t c = 0; c < nrhs; ++c)   // the levels of one right-hand side reuse the scratch of the previous one (stream order)
        if ((rc = solve_level<T>(h, n, dl, d, du, B + c * ldb, X + c * ldx, (T *)scr, info))) return rc;
    return 0;
}

}  // namespace

extern "C" {
int lbm_dtridiag_solve(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du, const double *B, int64_t ldb,
                       int64_t nrhs, double *X, int64_t ldx, int *info_dev) {
    return tridiag_solve_impl<double>(h, n, dl, d, du, B, ldb, nrhs, X, ldx, info_dev);
}
int lbm_stridiag_solve(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du, const float *B, int64_t ldb,
                       int64_t nrhs, float *X, int64_t ldx, int *info_dev) {
    return tridiag_solve_impl<float>(h, n, dl, d, du, B, ldb, nrhs, X, ldx, info_dev);
}
int lbm_dbidiag_solve(lbm_handle_t h, char uplo, int64_t n, const double *dv, const double *ev, const double *B, int64_t ldb,
                      int64_t nrhs, double *X, int64_t ldx, int *info_dev) {
    if (!h) return -1;
    if (uplo == 'U' || uplo == 'u') return tridiag_solve_impl<double>(h, n, nullptr, dv, ev, B, ldb, nrhs, X, ldx, info_dev);
    if (uplo == 'L' || uplo == 'l') return tridiag_solve_impl<double>(h, n, ev, dv, nullptr, B, ldb, nrhs, X, ldx, info_dev);
    return lbm_fail_arg(h, 2, "uplo must be 'U' or 'L'");
}
int lbm_sbidiag_solve(lbm_handle_t h, char uplo, int64_t n, const float *dv, const float *ev, const float *B, int64_t ldb,
                      int64_t nrhs, float *X, int64_t ldx, int *info_dev) {
    if (!h) return -1;
    if (uplo == 'U' || uplo == 'u') return tridiag_solve_impl<float>(h, n, nullptr, dv, ev, B, ldb, nrhs, X, ldx, info_dev);
    if (uplo == 'L' || uplo == 'l') return tridiag_solve_impl<float>(h, n, ev, dv, nullptr, B, ldb, nrhs, X, ldx, info_dev);
    return lbm_fail_arg(h, 2, "uplo must be 'U' or 'L'");
}
}
