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

// AB[(ku + i - j) + lda*j] for every band row r = ku + i - j in [0, kl+ku], column j in [0, n)
template <typename T>
__global__ void __launch_bounds__(SP_NT) tridiag_pack_kernel(int64_t n, const T *__restrict__ dl, const T *__restrict__ d,
                                                             const T *__restrict__ du, int kl, int ku, T *__restrict__ AB,
                                                             int64_t lda) {
    const int rows = kl + ku + 1;
    const int64_t total = (int64_t)rows * n;
    const int64_t stride = (int64_t)gridDim.x * SP_NT;
    for (int64_t e = (int64_t)blockIdx.x * SP_NT + threadIdx.x; e < total; e += stride) {
        const int64_t j = e / rows;
        const int r = (int)(e - j * rows);
        const int k = ku - r;   // diagonal index j - i
        T v = (T)0;
        if (k == 0) v = d[j];
        else if (k == 1) { if (du && j >= 1) v = du[j - 1]; }
        else if (k == -1) { if (dl && j + 1 < n) v = dl[j]; }
        AB[r + lda * j] = v;
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
    tridiag_pack_kernel<T><<<sp_grid<T>(h, (kl + ku + 1) * n), SP_NT, 0, h->stream>>>(n, dl, d, du, (int)kl, (int)ku, AB, lda);
    LBM_LAUNCH_CHECK(h, "tridiag_pack_kernel");
    return 0;
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
}
