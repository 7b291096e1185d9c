// special.cu -- `+` / `-` among Diagonal / Bidiagonal / Tridiagonal / SymTridiagonal / UniformScaling
// operands and their conversion to BandedMatrix storage (second half).
//
//   lbm_?tridiag_axpby : (odl, od, odu) <- alpha*(adl, ad, adu) + beta*(bdl, bd, bdu) + gamma*I
//       one launch for the three diagonals of every method of /root/reference/src/special.jl:75-222
//       (an absent diagonal is a NULL pointer: Diagonal has only `d`, an upper Bidiagonal only `d`,`du`, ...;
//        SymTridiagonal passes its `ev` as both dl and du; `gamma` is the UniformScaling term of :180-222).
//   lbm_?tridiag_pack  : the (kl,ku) diagonal-major array of the same matrix, i.e. `BandedMatrix(A, (kl,ku))`
//       for a structured operand (element definitions src/tridiag.jl:301-314, :471-484, src/bidiag.jl:111-124),
//       which is what routes structured x structured products (test/test_special.jl:154-160) to lbm_?gbmm.
//
// Both are pure streaming kernels (every operand read once, every result written once; algorithmic bytes =
// w * (elements read + elements written)), persistent grids of SMs x 8 CTAs, 128-bit accesses when the three
// pointers of a diagonal are 16-byte aligned.
#include "lbm_common.cuh"

namespace {

constexpr int SP_NT = 256;

template <typename T>
struct diag_job {
    const T *a, *b;
    T *o;
    int64_t len;
    T gamma;
    int vec;   // 1: a, b, o all 16-byte aligned (or absent)
};

template <typename T>
__device__ __forceinline__ T axpby_one(bool ha, bool hb, T alpha, T av, T beta, T bv, T gamma, bool hg) {
    T r = (T)0;
    if (ha) r = alpha * av;
    if (hb) r = ha ? fma(beta, bv, r) : beta * bv;
    if (hg) r += gamma;
    return r;
}

// blockIdx.y selects the diagonal (0 sub, 1 main, 2 super); each is a flat grid-stride stream
template <typename T>
__global__ void __launch_bounds__(SP_NT) tridiag_axpby_kernel(diag_job<T> j0, diag_job<T> j1, diag_job<T> j2, T alpha, T beta) {
    using V = typename lbm::vec16<T>::type;
    constexpr int EPV = lbm::vec16<T>::N;
    const diag_job<T> J = blockIdx.y == 0 ? j0 : (blockIdx.y == 1 ? j1 : j2);
    if (!J.o || J.len <= 0) return;
    const bool ha = J.a != nullptr, hb = J.b != nullptr, hg = J.gamma != (T)0;
    const int64_t tid = (int64_t)blockIdx.x * SP_NT + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * SP_NT;
    int64_t body = 0;
    if (J.vec) {
        const int64_t nv = J.len / EPV;
        body = nv * EPV;
        const V *av = reinterpret_cast<const V *>(J.a);
        const V *bv = reinterpret_cast<const V *>(J.b);
        V *ov = reinterpret_cast<V *>(J.o);
        for (int64_t v = tid; v < nv; v += stride) {
            V x, y, r;
            if (ha) x = lbm::ld_stream(av + v);
            if (hb) y = lbm::ld_stream(bv + v);
            T *xe = reinterpret_cast<T *>(&x), *ye = reinterpret_cast<T *>(&y), *re = reinterpret_cast<T *>(&r);
#pragma unroll
            for (int e = 0; e < EPV; ++e) re[e] = axpby_one<T>(ha, hb, alpha, ha ? xe[e] : (T)0, beta, hb ? ye[e] : (T)0, J.gamma, hg);
            lbm::st_stream(ov + v, r);
        }
    }
    for (int64_t i = body + tid; i < J.len; i += stride)
        J.o[i] = axpby_one<T>(ha, hb, alpha, ha ? J.a[i] : (T)0, beta, hb ? J.b[i] : (T)0, J.gamma, hg);
}

// AB[(ku + i - j) + lda*j] for every band row r = ku + i - j in [0, kl+ku], column j in [0, n).
// A CTA walks tiles of SP_PACK_TC columns; inside a tile the flat element index is 32-bit (one cheap division).
constexpr int SP_PACK_TC = 2048;
template <typename T>
__global__ void __launch_bounds__(SP_NT) tridiag_pack_kernel(int64_t n, const T *__restrict__ dl, const T *__restrict__ d,
                                                             const T *__restrict__ du, int kl, int ku, T *__restrict__ AB,
                                                             int64_t lda) {
    const unsigned rows = (unsigned)(kl + ku + 1);
    const int64_t ntiles = (n + SP_PACK_TC - 1) / SP_PACK_TC;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t j0 = t * SP_PACK_TC;
        const unsigned cols = (unsigned)((n - j0) < SP_PACK_TC ? (n - j0) : SP_PACK_TC);
        const unsigned total = rows * cols;
        // flat index e = c*rows + r advanced incrementally (no division in the loop)
        const unsigned dc = SP_NT / rows, dr = SP_NT - dc * rows;
        unsigned c = threadIdx.x / rows, r = threadIdx.x - c * rows;
        for (unsigned e = threadIdx.x; e < total; e += SP_NT) {
            const int64_t j = j0 + c;
            const int k = ku - (int)r;   // diagonal index j - i
            T v = (T)0;
            if (k == 0) v = d[j];
            else if (k == 1) { if (du && j >= 1) v = du[j - 1]; }
            else if (k == -1) { if (dl && j + 1 < n) v = dl[j]; }
            AB[r + lda * j] = v;
            r += dr;
            c += dc;
            if (r >= rows) { r -= rows; ++c; }
        }
    }
}

// C = A*B for two structured operands, written straight into the (la+lb, ua+ub) diagonal-major array
// (test/test_special.jl:154-160, test/test_bidiag.jl:262-267 all-pairs products).  Column j of C:
//   C[j-2,j] = au[j-2] bu[j-1]
//   C[j-1,j] = ad[j-1] bu[j-1] + au[j-1] bd[j]
//   C[j  ,j] = al[j-1] bu[j-1] + ad[j] bd[j] + au[j] bl[j]
//   C[j+1,j] = al[j] bd[j] + ad[j+1] bl[j]
//   C[j+2,j] = al[j+1] bl[j]
// (al/ad/au = A's sub/main/super diagonal indexed by COLUMN for al, ad and by ROW for au, as the reference stores them;
//  sums run over k ascending like the banded oracle).  A thread owns a column; the tile is staged in shared memory
// so the band array is written as one contiguous, fully coalesced range.
constexpr int SP_MM_TC = 512;
template <typename T, bool AL, bool AU, bool BL, bool BU>
__global__ void __launch_bounds__(SP_NT) tridiag_mm_kernel(int64_t n, const T *__restrict__ al, const T *__restrict__ ad,
                                                           const T *__restrict__ au, const T *__restrict__ bl,
                                                           const T *__restrict__ bd, const T *__restrict__ bu,
                                                           T *__restrict__ C, int64_t ldc) {
    constexpr int LC = (AL ? 1 : 0) + (BL ? 1 : 0), UC = (AU ? 1 : 0) + (BU ? 1 : 0), ROWS = LC + UC + 1;
    __shared__ T tile[SP_MM_TC * ROWS];
    const int64_t ntiles = (n + SP_MM_TC - 1) / SP_MM_TC;
    const int64_t m = n - 1;   // length of the off-diagonals
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t j0 = t * SP_MM_TC;
        const int cols = (int)((n - j0) < SP_MM_TC ? (n - j0) : SP_MM_TC);
        for (int c = threadIdx.x; c < cols; c += SP_NT) {
            const int64_t j = j0 + c;
            auto ld = [](const T *p, int64_t i, int64_t len) { return (i >= 0 && i < len) ? p[i] : (T)0; };
            const T bdj = bd[j];
            const T buj = BU ? ld(bu, j - 1, m) : (T)0;      // B[j-1,j]
            const T blj = BL ? ld(bl, j, m) : (T)0;          // B[j+1,j]
            T v[5] = {(T)0, (T)0, (T)0, (T)0, (T)0};         // C[j-2..j+2, j]
            if (AU && BU) v[0] = ld(au, j - 2, m) * buj;
            if (BU) v[1] = ld(ad, j - 1, n) * buj;
            if (AU) v[1] = BU ? fma(ld(au, j - 1, m), bdj, v[1]) : ld(au, j - 1, m) * bdj;
            if (AL && BU) v[2] = ld(al, j - 1, m) * buj;
            v[2] = (AL && BU) ? fma(ad[j], bdj, v[2]) : ad[j] * bdj;
            if (AU && BL) v[2] = fma(ld(au, j, m), blj, v[2]);
            if (AL) v[3] = ld(al, j, m) * bdj;
            if (BL) v[3] = AL ? fma(ld(ad, j + 1, n), blj, v[3]) : ld(ad, j + 1, n) * blj;
            if (AL && BL) v[4] = ld(al, j + 1, m) * blj;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const int off = r - UC;                      // i - j
                const int64_t i = j + off;
                tile[c * ROWS + r] = (i >= 0 && i < n) ? v[off + 2] : (T)0;
            }
        }
        __syncthreads();
        const int total = cols * ROWS;
        if (ldc == ROWS) {
            T *dst = C + j0 * ldc;
            for (int e = threadIdx.x; e < total; e += SP_NT) dst[e] = tile[e];
        } else {
            for (int e = threadIdx.x; e < total; e += SP_NT) {
                const int c = e / ROWS, r = e - c * ROWS;
                C[r + ldc * (j0 + c)] = tile[e];
            }
        }
        __syncthreads();
    }
}

template <typename T>
int sp_grid(lbm_handle_t h, int64_t work) {
    int64_t g = (work + SP_NT - 1) / SP_NT;
    const int64_t cap = (int64_t)h->sm_count * 8;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

template <typename T>
int tridiag_axpby_impl(lbm_handle_t h, int64_t n, T alpha, const T *adl, const T *ad, const T *adu, T beta, const T *bdl,
                       const T *bd, const T *bdu, T gamma, T *odl, T *od, T *odu) {
    if (!h) return -1;
    if (n < 0) return lbm_fail_arg(h, 2, "n < 0");
    if (n == 0) return 0;
    if (!od && !odl && !odu) return lbm_fail_arg(h, 13, "no output diagonal");
    lbm_device_guard g(h->device);
    auto job = [](const T *a, const T *b, T *o, int64_t len, T gm) {
        diag_job<T> J;
        J.a = a; J.b = b; J.o = o; J.len = o ? len : 0; J.gamma = gm;
        J.vec = (!a || lbm_aligned16(a)) && (!b || lbm_aligned16(b)) && lbm_aligned16(o);
        return J;
    };
    const diag_job<T> j0 = job(adl, bdl, odl, n - 1, (T)0), j1 = job(ad, bd, od, n, gamma), j2 = job(adu, bdu, odu, n - 1, (T)0);
    constexpr int EPV = lbm::vec16<T>::N;
    dim3 grid(sp_grid<T>(h, (n + EPV - 1) / EPV), 3);
    tridiag_axpby_kernel<T><<<grid, SP_NT, 0, h->stream>>>(j0, j1, j2, alpha, beta);
    LBM_LAUNCH_CHECK(h, "tridiag_axpby_kernel");
    return 0;
}

template <typename T>
int tridiag_pack_impl(lbm_handle_t h, int64_t n, const T *dl, const T *d, const T *du, int64_t kl, int64_t ku, T *AB,
                      int64_t lda) {
    if (!h) return -1;
    if (n < 0) return lbm_fail_arg(h, 2, "n < 0");
    if (kl < 0 || kl > 1024) return lbm_fail_arg(h, 6, "kl out of range");
    if (ku < 0 || ku > 1024) return lbm_fail_arg(h, 7, "ku out of range");
    if (lda < kl + ku + 1) return lbm_fail_arg(h, 9, "lda < kl+ku+1");
    if (n == 0) return 0;
    if (!d) return lbm_fail_arg(h, 4, "d == NULL");
    if (!AB) return lbm_fail_arg(h, 8, "AB == NULL");
    if (dl && kl < 1) return lbm_fail_arg(h, 6, "kl == 0 but a subdiagonal was passed");
    if (du && ku < 1) return lbm_fail_arg(h, 7, "ku == 0 but a superdiagonal was passed");
    lbm_device_guard g(h->device);
    tridiag_pack_kernel<T><<<sp_grid<T>(h, ((n + SP_PACK_TC - 1) / SP_PACK_TC) * SP_NT), SP_NT, 0, h->stream>>>(n, dl, d, du, (int)kl, (int)ku, AB, lda);
    LBM_LAUNCH_CHECK(h, "tridiag_pack_kernel");
    return 0;
}

template <typename T, bool AL, bool AU>
int tridiag_mm_launch(lbm_handle_t h, int64_t n, const T *al, const T *ad, const T *au, const T *bl, const T *bd, const T *bu,
                      T *C, int64_t ldc) {
    const int grid = sp_grid<T>(h, ((n + SP_MM_TC - 1) / SP_MM_TC) * SP_NT);
#define LBM_MM(BL_, BU_) tridiag_mm_kernel<T, AL, AU, BL_, BU_><<<grid, SP_NT, 0, h->stream>>>(n, al, ad, au, bl, bd, bu, C, ldc)
    if (bl && bu) LBM_MM(true, true);
    else if (bl) LBM_MM(true, false);
    else if (bu) LBM_MM(false, true);
    else LBM_MM(false, false);
#undef LBM_MM
    LBM_LAUNCH_CHECK(h, "tridiag_mm_kernel");
    return 0;
}

template <typename T>
int tridiag_mm_impl(lbm_handle_t h, int64_t n, const T *adl, const T *ad, const T *adu, const T *bdl, const T *bd, const T *bdu,
                    T *C, int64_t ldc) {
    if (!h) return -1;
    if (n < 0) return lbm_fail_arg(h, 2, "n < 0");
    const int64_t rows = (adl ? 1 : 0) + (bdl ? 1 : 0) + (adu ? 1 : 0) + (bdu ? 1 : 0) + 1;
    if (ldc < rows) return lbm_fail_arg(h, 10, "ldc < la+lb+ua+ub+1");
    if (n == 0) return 0;
    if (!ad) return lbm_fail_arg(h, 4, "ad == NULL");
    if (!bd) return lbm_fail_arg(h, 7, "bd == NULL");
    if (!C) return lbm_fail_arg(h, 9, "C == NULL");
    lbm_device_guard g(h->device);
    if (adl && adu) return tridiag_mm_launch<T, true, true>(h, n, adl, ad, adu, bdl, bd, bdu, C, ldc);
    if (adl) return tridiag_mm_launch<T, true, false>(h, n, adl, ad, adu, bdl, bd, bdu, C, ldc);
    if (adu) return tridiag_mm_launch<T, false, true>(h, n, adl, ad, adu, bdl, bd, bdu, C, ldc);
    return tridiag_mm_launch<T, false, false>(h, n, adl, ad, adu, bdl, bd, bdu, C, ldc);
}

}  // namespace

extern "C" {
int lbm_dtridiag_axpby(lbm_handle_t h, int64_t n, double alpha, const double *adl, const double *ad, const double *adu,
                       double beta, const double *bdl, const double *bd, const double *bdu, double gamma, double *odl,
                       double *od, double *odu) {
    return tridiag_axpby_impl<double>(h, n, alpha, adl, ad, adu, beta, bdl, bd, bdu, gamma, odl, od, odu);
}
int lbm_stridiag_axpby(lbm_handle_t h, int64_t n, float alpha, const float *adl, const float *ad, const float *adu, float beta,
                       const float *bdl, const float *bd, const float *bdu, float gamma, float *odl, float *od, float *odu) {
    return tridiag_axpby_impl<float>(h, n, alpha, adl, ad, adu, beta, bdl, bd, bdu, gamma, odl, od, odu);
}
int lbm_dtridiag_pack(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du, int64_t kl, int64_t ku,
                      double *AB, int64_t lda) {
    return tridiag_pack_impl<double>(h, n, dl, d, du, kl, ku, AB, lda);
}
int lbm_stridiag_pack(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du, int64_t kl, int64_t ku,
                      float *AB, int64_t lda) {
    return tridiag_pack_impl<float>(h, n, dl, d, du, kl, ku, AB, lda);
}
int lbm_dtridiag_mm(lbm_handle_t h, int64_t n, const double *adl, const double *ad, const double *adu, const double *bdl,
                    const double *bd, const double *bdu, double *C, int64_t ldc) {
    return tridiag_mm_impl<double>(h, n, adl, ad, adu, bdl, bd, bdu, C, ldc);
}
int lbm_stridiag_mm(lbm_handle_t h, int64_t n, const float *adl, const float *ad, const float *adu, const float *bdl,
                    const float *bd, const float *bdu, float *C, int64_t ldc) {
    return tridiag_mm_impl<float>(h, n, adl, ad, adu, bdl, bd, bdu, C, ldc);
}
}
