/*
 * lbm_b200.h -- C ABI of liblbm_b200.so: the B200 (sm_100a) banded mul!/materialize
 * hot path behind LazyBandedMatrices.jl's API.
 *
 * What this boundary replaces.  In the reference the path ends in the only FFI it
 * has: BandedMatrices.jl `gbmv!` -> `ccall((@blasfunc(dgbmv_), libblastrampoline), ...)`
 * with the Fortran signature (trans,m,n,kl,ku,alpha,A,lda,x,incx,beta,y,incy)
 * (un-vendored dependency, /root/reference/Project.toml:7; reached from the layout
 * traits the reference declares at src/tridiag.jl:685-688, src/bidiag.jl:438-440,
 * src/LazyBandedMatrices.jl:14-21).  Every entry point below is what a Julia host
 * shim binds with `ccall` in place of that BLAS call (see INTEGRATION.md and
 * lazybandedmatrices.jl_b200/julia/LBMB200.jl); argument meaning follows BLAS/LAPACK.
 *
 * Conventions
 *  - Plain C: pointers + int64 sizes, no C++/torch types.  All matrix/vector pointers
 *    are DEVICE pointers on the handle's GPU unless the function name ends in `_host`.
 *  - Band storage = BandedMatrices `(kl+ku+1) x n` column-major "AB" format:
 *        A[i,j] = data[(ku + i - j) + lda*j]   (0-based), -ku <= i-j <= kl.
 *    Out-of-matrix slots are never read (masked by INDEX, never by value).
 *  - Return value (LAPACK `info` style): 0 ok; -k = k-th argument invalid (the host
 *    shim maps it to DimensionMismatch/ArgumentError, mirroring
 *    /root/reference/src/bidiag.jl:363-377 and test/test_tridiag.jl:272-279);
 *    >0 = CUDA error (cudaError_t), text via lbm_last_error_string().
 *  - Calls ENQUEUE on the handle's stream and return; lbm_sync() waits.  `_host`
 *    variants copy in, compute, copy out and synchronise before returning.
 *  - beta == 0 means y/C is not read (BLAS convention; NaNs do not propagate).
 *  - A handle is bound to one device and is not thread-safe.
 *  - There is NO CPU fallback: without a CUDA device lbm_create() fails.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lbm_handle_s *lbm_handle_t;

/* ---- handle ---------------------------------------------------------------------- */
int lbm_version(void);                                   /* 10000*major + 100*minor + patch */
int lbm_create(lbm_handle_t *out, int device);           /* owns a non-blocking stream */
int lbm_destroy(lbm_handle_t h);
int lbm_set_stream(lbm_handle_t h, void *cuda_stream);   /* borrow caller's cudaStream_t; NULL = CUDA's legacy default stream */
int lbm_use_own_stream(lbm_handle_t h);                  /* back to the handle's private non-blocking stream */
int lbm_sync(lbm_handle_t h);
const char *lbm_last_error_string(lbm_handle_t h);       /* h may be NULL: last create error */
int64_t lbm_launch_count(lbm_handle_t h);                /* kernels launched through h so far */
int lbm_set_tuning(lbm_handle_t h, const char *key, int64_t value); /* kernel selection knobs */
/* Keys (0 = automatic everywhere; unknown key: -2).  They select among kernels / configurations that all compute the same result:
 *   gbmv_tile, gbmv_stages, gbmv_ctas_per_sm, gbmv_force (1 direct, 2 tile ring, 3 wide-N, 4 warp-per-output T), gbmvw_chunk, gbmvw_stages,
 *   gbmvw_ctas_per_sm; tri_tile, tri_stages, tri_ctas_per_sm, tri_force_scalar, tri_force_shuffle;
 *   kron_tr (256/512/1024 rows per strip), kron_tc (grid columns per chunk), kron_stages, kron_ctas_per_sm, kron_force_tile;
 *   gbmm_force (1 diagonal-wise, 2 FP64 DMMA, 3 register kernel, 4 register kernel with round-1 zero-padded staging),
 *   gbmm_tile, gbmm_variant (DMMA kernel variants; 30-32 = warp-autonomous); chain_tile, chain_force_generic, chain_variant;
 *   exchange_mode (1 = ncclSend/Recv instead of peer mailboxes), exchange_fuse (2 = two-kernel exchange instead of the
 *   exchange fused into the compute launch), halo_timeout_ms (wall-clock bound of a ghost wait, default 30000),
 *   pdl (1 = programmatic dependent launch of the ring kernels; measured slower, off). */
int64_t lbm_get_tuning(lbm_handle_t h, const char *key);

/* ---- memory helpers (for hosts without their own CUDA allocator, e.g. the Julia shim) - */
int lbm_malloc(lbm_handle_t h, void **dptr, size_t bytes);
int lbm_free(lbm_handle_t h, void *dptr);
int lbm_host_alloc(lbm_handle_t h, void **hptr, size_t bytes);   /* pinned */
int lbm_host_free(lbm_handle_t h, void *hptr);
int lbm_memcpy_h2d(lbm_handle_t h, void *dst, const void *src, size_t bytes);  /* async on stream */
int lbm_memcpy_d2h(lbm_handle_t h, void *dst, const void *src, size_t bytes);
int lbm_memcpy_d2d(lbm_handle_t h, void *dst, const void *src, size_t bytes);
int lbm_memset0(lbm_handle_t h, void *dst, size_t bytes);

/* ---- synthetic data: counter-based U[0,1) (same stream as oracle/) ---- */
int lbm_dfill_uniform(lbm_handle_t h, uint64_t seed, int64_t offset, int64_t count, double *dst);
int lbm_sfill_uniform(lbm_handle_t h, uint64_t seed, int64_t offset, int64_t count, float *dst);

/* ---- banded matrix * vector:  y <- alpha*op(A)*x + beta*y  --------------------------- */
/* Replaces BandedMatrices.gbmv! -> BLAS ?gbmv_; trans in {'N','T','C'}. */
int lbm_dgbmv(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku,
              double alpha, const double *A, int64_t lda, const double *x, double beta, double *y);
int lbm_sgbmv(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku,
              float alpha, const float *A, int64_t lda, const float *x, float beta, float *y);
/* banded * dense, `mul!(C, A::BandedMatrix, B::Matrix, alpha, beta)` ([up] BandedMatrices; exercised by the reference
 * at test/test_tridiag.jl:253 for the structured kinds): Y[:,c] <- alpha*op(A)*X[:,c] + beta*Y[:,c], c < nrhs, X / Y
 * column-major with leading dimensions ldx / ldy.  Narrow bands share ONE pass over A between up to 8 columns. */
int lbm_dgbmv_multi(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, double alpha,
                    const double *A, int64_t lda, const double *X, int64_t ldx, double beta, double *Y, int64_t ldy,
                    int64_t nrhs);
int lbm_sgbmv_multi(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, float alpha,
                    const float *A, int64_t lda, const float *X, int64_t ldx, float beta, float *Y, int64_t ldy,
                    int64_t nrhs);
/* fused broadcast sum, `mul!(y, BroadcastArray(+/-, A1, A2), x)` (BroadcastBandedLayout, reference
 * src/LazyBandedMatrices.jl:20; test/test_misc.jl:21-40, test/test_blockkron.jl:303,345):
 * y <- alpha1*A1*x + alpha2*A2*x + beta*y without forming A1 +- A2 and with ONE pass over x and y. */
int lbm_dgbmv2(lbm_handle_t h, int64_t m, int64_t n, double alpha1, int64_t kl1, int64_t ku1, const double *A1, int64_t lda1,
               double alpha2, int64_t kl2, int64_t ku2, const double *A2, int64_t lda2, const double *x, double beta,
               double *y);
int lbm_sgbmv2(lbm_handle_t h, int64_t m, int64_t n, float alpha1, int64_t kl1, int64_t ku1, const float *A1, int64_t lda1,
               float alpha2, int64_t kl2, int64_t ku2, const float *A2, int64_t lda2, const float *x, float beta, float *y);
/* same, HOST pointers (pageable or pinned): H2D(A,x[,y]) + kernel + D2H(y) + sync. */
int lbm_dgbmv_host(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku,
                   double alpha, const double *A, int64_t lda, const double *x, double beta, double *y);
int lbm_sgbmv_host(lbm_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku,
                   float alpha, const float *A, int64_t lda, const float *x, float beta, float *y);

/* ---- Tridiagonal / SymTridiagonal / Bidiagonal:  Y <- alpha*A*X + beta*Y ------------- */
/* A = Tridiagonal(dl,d,du), n x n (reference src/tridiag.jl:327-341, elements :471-484).
 * SymTridiagonal(dv,ev) passes dl == du == ev (src/tridiag.jl:6-16, :301-314; only
 * ev[0..n-2] is read).  Bidiagonal(dv,ev,'U') passes dl == NULL, 'L' passes du == NULL
 * (src/bidiag.jl:111-124).  X, Y are n x nrhs column-major with leading dims ldx, ldy. */
int lbm_dtridiag_mv(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du,
                    double alpha, const double *X, int64_t ldx, int64_t nrhs,
                    double beta, double *Y, int64_t ldy);
int lbm_stridiag_mv(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du,
                    float alpha, const float *X, int64_t ldx, int64_t nrhs,
                    float beta, float *Y, int64_t ldy);
/* mat-vec restricted to rows [row_begin,row_end): y[i-row_begin] <- alpha*(A x)[i] + beta*y[i-row_begin].
 * Used by row shards: the rank applies its ghost-extended square system and keeps the owned rows. */
int lbm_dtridiag_mv_rows(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du,
                         double alpha, const double *x, double beta, double *y, int64_t row_begin, int64_t row_end);
int lbm_stridiag_mv_rows(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du,
                         float alpha, const float *x, float beta, float *y, int64_t row_begin, int64_t row_end);
int lbm_dbidiag_mv(lbm_handle_t h, char uplo, int64_t n, const double *dv, const double *ev,
                   double alpha, const double *X, int64_t ldx, int64_t nrhs,
                   double beta, double *Y, int64_t ldy);
int lbm_sbidiag_mv(lbm_handle_t h, char uplo, int64_t n, const float *dv, const float *ev,
                   float alpha, const float *X, int64_t ldx, int64_t nrhs,
                   float beta, float *Y, int64_t ldy);

/* ---- banded * banded -> banded:  C <- alpha*A*B + beta*C ------------------------------ */
/* ---- single-pass reductions over the structured kinds --------- */
/* dot(x, A, y) = x' A y for A = Tridiagonal(dl,d,du) (reference src/tridiag.jl:665-682); SymTridiagonal passes
 * dl == du == ev, Bidiagonal passes NULL for the missing diagonal.  *result is a DEVICE scalar, written by a
 * deterministic two-stage reduction (no atomics).  n == 0 -> 0. */
int lbm_dtridiag_dot(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du, const double *x,
                     const double *y, double *result);
int lbm_stridiag_dot(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du, const float *x,
                     const float *y, float *result);
/* Base._sum(A, dims) (reference src/tridiag.jl:594-660, src/bidiag.jl:396-435): dims = 1 -> out[j] = sum_i A[i,j]
 * (n values), dims = 2 -> out[i] = sum_j A[i,j] (n values), dims = 0 -> the scalar sum(A) (Colon). */
int lbm_dtridiag_sum(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du, int dims, double *out);
int lbm_stridiag_sum(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du, int dims, float *out);

/* Materialises ApplyMatrix(*, A, B) (reference README.md:8-26; [up] BandedMatrices gbmm!).
 * A m x k (kla,kua), B k x n (klb,kub), C m x n stored with band (klc,kuc) which must
 * cover the product band clipped to the matrix: klc >= min(kla+klb, m-1),
 * kuc >= min(kua+kub, n-1).  Out-of-matrix slots of C are written 0 when beta == 0.
 * lbm_sgbmm with both factors >= 33 stored diagonals and a tight ldc computes in Float64 on the tensor path (operands widened
 * into handle-owned scratch, 8 bytes per stored element of A, B and C; the diagonal-wise Float32 kernel when that is not free). */
int lbm_dgbmm(lbm_handle_t h, int64_t m, int64_t n, int64_t k, double alpha,
              int64_t kla, int64_t kua, const double *A, int64_t lda,
              int64_t klb, int64_t kub, const double *B, int64_t ldb, double beta,
              int64_t klc, int64_t kuc, double *C, int64_t ldc);
int lbm_sgbmm(lbm_handle_t h, int64_t m, int64_t n, int64_t k, float alpha,
              int64_t kla, int64_t kua, const float *A, int64_t lda,
              int64_t klb, int64_t kub, const float *B, int64_t ldb, float beta,
              int64_t klc, int64_t kuc, float *C, int64_t ldc);

/* ---- batched lazy chain  D_b = (A_b*B_b)*C_b,  b = 0..batch-1 -------------------------- */
/* Square n x n factors with tight leading dimension kl[i]+ku[i]+1; item b of factor F
 * starts at F + b*strideF (elements).  D has band (sum kl, sum ku), tight ld.  The
 * intermediate band never touches HBM. */
int lbm_sgbmm_chain3_batched(lbm_handle_t h, int64_t batch, int64_t n,
                             const int64_t *kl, const int64_t *ku,
                             const float *A, int64_t strideA, const float *B, int64_t strideB,
                             const float *C, int64_t strideC, float *D, int64_t strideD);
int lbm_dgbmm_chain3_batched(lbm_handle_t h, int64_t batch, int64_t n,
                             const int64_t *kl, const int64_t *ku,
                             const double *A, int64_t strideA, const double *B, int64_t strideB,
                             const double *C, int64_t strideC, double *D, int64_t strideD);

/* ---- block-Kronecker operators applied without forming them -------------------------- */
/* y <- alpha*(kron(A,I_n2) + kron(I_n1,B))*x + beta*y = alpha*vec(X*A^T + B*X) + beta*y,
 * x = vec(X), X n2 x n1 column-major; A n1 x n1 (kla,kua), B n2 x n2 (klb,kub) banded.
 * This is BlockKron(A,I)+BlockKron(I,B) -- the 2-D Laplacian of
 * /root/reference/test/test_blockkron.jl:296-306 -- via the identity the reference uses at
 * src/blockkron.jl:272-278,:440. */
int lbm_dkronsum2d_mv(lbm_handle_t h, int64_t n1, int64_t n2,
                      int64_t kla, int64_t kua, const double *A, int64_t lda,
                      int64_t klb, int64_t kub, const double *B, int64_t ldb,
                      double alpha, const double *x, double beta, double *y);
int lbm_skronsum2d_mv(lbm_handle_t h, int64_t n1, int64_t n2,
                      int64_t kla, int64_t kua, const float *A, int64_t lda,
                      int64_t klb, int64_t kub, const float *B, int64_t ldb,
                      float alpha, const float *x, float beta, float *y);
/* y <- alpha*kron(A,B)*x + beta*y = alpha*vec(B*X*A^T) + beta*y  (BlockKron(A,B)*x;
 * work = n1*n2 elements of scratch owned by the handle). */
int lbm_dkron2d_mv(lbm_handle_t h, int64_t n1, int64_t n2,
                   int64_t kla, int64_t kua, const double *A, int64_t lda,
                   int64_t klb, int64_t kub, const double *B, int64_t ldb,
                   double alpha, const double *x, double beta, double *y);
int lbm_skron2d_mv(lbm_handle_t h, int64_t n1, int64_t n2,
                   int64_t kla, int64_t kua, const float *A, int64_t lda,
                   int64_t klb, int64_t kub, const float *B, int64_t ldb,
                   float alpha, const float *x, float beta, float *y);

/* y <- alpha*(A (x) B (x) C)*x + beta*y on a full n x n x n tensor (k fastest): the three mode
 * products of /root/reference/src/blockkron.jl:441-450 (A on dim 3, B on dim 2, C on dim 1), i.e.
 * the arithmetic of KronTrav(A,B,C)*DiagTrav(X) before the DiagTrav re-ordering. */
int lbm_dkron3d_mv(lbm_handle_t h, int64_t n,
                   int64_t kla, int64_t kua, const double *A, int64_t lda,
                   int64_t klb, int64_t kub, const double *B, int64_t ldb,
                   int64_t klc, int64_t kuc, const double *C, int64_t ldc,
                   double alpha, const double *x, double beta, double *y);
int lbm_skron3d_mv(lbm_handle_t h, int64_t n,
                   int64_t kla, int64_t kua, const float *A, int64_t lda,
                   int64_t klb, int64_t kub, const float *B, int64_t ldb,
                   int64_t klc, int64_t kuc, const float *C, int64_t ldc,
                   float alpha, const float *x, float beta, float *y);

/* DiagTrav ordering on the device (/root/reference/src/blockkron.jl:14-24, :101-133; zero_bottomright :27-64; InvDiagTrav
 * :226-256): X is the m x n matrix (ndim = 2) or n x n x n tensor (ndim = 3, m == n) in column-major order, v the block
 * vector of its anti-diagonals; entries below the last anti-diagonal block (the "bottom right") do not belong to v and
 * are written as zeros by lbm_?invdiagtrav / lbm_?zero_bottomright (in place).  lbm_diagtrav_length = length(v). */
int lbm_ddiagtrav(lbm_handle_t h, int ndim, int64_t m, int64_t n, const double *X, double *v);
int lbm_sdiagtrav(lbm_handle_t h, int ndim, int64_t m, int64_t n, const float *X, float *v);
int lbm_dinvdiagtrav(lbm_handle_t h, int ndim, int64_t m, int64_t n, const double *v, double *X);
int lbm_sinvdiagtrav(lbm_handle_t h, int ndim, int64_t m, int64_t n, const float *v, float *X);
int lbm_dzero_bottomright(lbm_handle_t h, int ndim, int64_t m, int64_t n, double *X);
int lbm_szero_bottomright(lbm_handle_t h, int ndim, int64_t m, int64_t n, float *X);
int64_t lbm_diagtrav_length(int ndim, int64_t m, int64_t n);

/* ---- lazy product applied to a vector: y <- alpha*F_0*(F_1*(...F_{nf-1}*x)) + beta*y ------------------ */
/* mul!(y, ApplyMatrix(*, F_0, ..., F_{nf-1}), x, alpha, beta): LazyArrays' right-to-left Mul flattening (the rule the
 * reference relies on at /root/reference/src/blockkron.jl:208-214, pinned at test/test_blockkron.jl:345-348) as ONE call:
 * factor f is m[f] x n[f] with bandwidths (kl[f], ku[f]), data A[f], leading dimension lda[f]; n[f] == m[f+1].
 * nf launches on the handle's stream; the nf-1 temporaries live in the handle's scratch (no allocation per call). */
int lbm_dgbmv_chain(lbm_handle_t h, int nf, const int64_t *m, const int64_t *n, const int64_t *kl, const int64_t *ku,
                    const double *const *A, const int64_t *lda, double alpha, const double *x, double beta, double *y);
int lbm_sgbmv_chain(lbm_handle_t h, int nf, const int64_t *m, const int64_t *n, const int64_t *kl, const int64_t *ku,
                    const float *const *A, const int64_t *lda, float alpha, const float *x, float beta, float *y);

/* ---- `+` / `-` among structured kinds, and their BandedMatrix storage ----------------------- */
/* (odl, od, odu) <- alpha*(adl, ad, adu) + beta*(bdl, bd, bdu) + gamma*I on the three diagonals of an n x n
 * operand in ONE launch: every method of /root/reference/src/special.jl:75-222 (Diagonal / Bidiagonal /
 * Tridiagonal / SymTridiagonal pairs, and the UniformScaling methods through gamma).  A diagonal a kind does not
 * have is NULL (sub/super have n-1 entries); an output pointer may be NULL when that diagonal is not wanted
 * (the reference re-uses the operand's array there, e.g. `Bidiagonal(newdv, A.ev, A.uplo)`, special.jl:85-88).
 * alpha, beta = +-1 reproduce Julia's `a + b`, `a - b`, `-b` bit for bit. */
int lbm_dtridiag_axpby(lbm_handle_t h, int64_t n, double alpha, const double *adl, const double *ad, const double *adu,
                       double beta, const double *bdl, const double *bd, const double *bdu, double gamma,
                       double *odl, double *od, double *odu);
int lbm_stridiag_axpby(lbm_handle_t h, int64_t n, float alpha, const float *adl, const float *ad, const float *adu,
                       float beta, const float *bdl, const float *bd, const float *bdu, float gamma,
                       float *odl, float *od, float *odu);
/* AB <- the (kl,ku) diagonal-major array of the structured operand (dl, d, du), zeros on the other band rows:
 * `BandedMatrix(A, (kl,ku))` with the element definitions of src/tridiag.jl:301-314, :471-484 and
 * src/bidiag.jl:111-124.  Routes structured x structured products (test/test_special.jl:154-160,
 * test/test_bidiag.jl:262-267) to lbm_?gbmm. */
int lbm_dtridiag_pack(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du,
                      int64_t kl, int64_t ku, double *AB, int64_t lda);
int lbm_stridiag_pack(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du,
                      int64_t kl, int64_t ku, float *AB, int64_t lda);

/* C <- A*B for two structured operands, written straight into the diagonal-major array of the product:
 * bandwidths (la+lb, ua+ub) with la = (adl != NULL), ua = (adu != NULL), ..., leading dimension ldc >= la+lb+ua+ub+1;
 * out-of-matrix slots are written as zeros.  One pass: each diagonal read once, the band written once
 * (the all-pairs products of test/test_special.jl:154-160 / test/test_bidiag.jl:262-267; in the reference they go
 * through ArrayLayouts' banded mul, one generic loop per column). */
int lbm_dtridiag_mm(lbm_handle_t h, int64_t n, const double *adl, const double *ad, const double *adu,
                    const double *bdl, const double *bd, const double *bdu, double *C, int64_t ldc);
int lbm_stridiag_mm(lbm_handle_t h, int64_t n, const float *adl, const float *ad, const float *adu,
                    const float *bdl, const float *bd, const float *bdu, float *C, int64_t ldc);

/* ---- solves with the structured kinds:  X <- A \ B ---------------------------------------------------------------- */
/* `A \ b` for Tridiagonal / SymTridiagonal (pass dl == du == ev) / Bidiagonal (a triangular solve: the missing
 * diagonal is NULL, or lbm_?bidiag_solve with uplo) -- SURVEY.md 8(f) rank 4; exercised by the reference at
 * /root/reference/test/test_tridiag.jl:345-363 (`A90\b90`), where `\` falls through ArrayLayouts to LinearAlgebra's
 * sequential tridiagonal LU.  Here: the partition method, recursively (csrc/solve.cu), 9n elements of HBM traffic.
 * B, X are n x nrhs column-major (ldb, ldx); X may alias B (in place).  NO PIVOTING (Thomas): exact for diagonally
 * dominant and symmetric positive definite operators.  info_dev (DEVICE int, may be NULL) is set to 1 when a zero pivot
 * was met (the result then holds Inf/NaN), 0 otherwise -- LAPACK's info > 0. */
int lbm_dtridiag_solve(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du,
                       const double *B, int64_t ldb, int64_t nrhs, double *X, int64_t ldx, int *info_dev);
int lbm_stridiag_solve(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du,
                       const float *B, int64_t ldb, int64_t nrhs, float *X, int64_t ldx, int *info_dev);
int lbm_dbidiag_solve(lbm_handle_t h, char uplo, int64_t n, const double *dv, const double *ev,
                      const double *B, int64_t ldb, int64_t nrhs, double *X, int64_t ldx, int *info_dev);
int lbm_sbidiag_solve(lbm_handle_t h, char uplo, int64_t n, const float *dv, const float *ev,
                      const float *B, int64_t ldb, int64_t nrhs, float *X, int64_t ldx, int *info_dev);

/* ---- row-sharded mul! across the GPUs of one box (one process per GPU) ------------------- */
/* Rank r of `world` owns rows [r0,r1) of a square n x n operator and stores the COLUMN slice
 * [c0,c1) = [max(0,r0-kl), min(n,r1+ku)) of the diagonal-major array (same lda); x lives in a
 * ghost-cell vector [left | owned | right] of length c1-c0.  lbm_shard_plan writes
 * {r0, r1, c0, c1, left, right, kl_loc, ku_loc} (pure integers, no GPU).  The communicator is
 * NCCL (dlopen'ed): rank 0 calls lbm_comm_unique_id, the host broadcasts the 128 bytes, every
 * rank calls lbm_comm_init.  lbm_?gbmv_sharded: y_own <- alpha*A[r0:r1,:]*x + beta*y_own.  With the peer mailboxes
 * connected (below) and overlap = 1 the ghost exchange is FUSED into the compute launch (csrc/halo.cuh): one kernel per
 * apply pushes this rank's boundary values over NVLink at its start and patches the ghost part of its last tiles' x
 * windows in shared memory -- the ghost CELLS of xbuf are then scratch (unspecified after the call).  Otherwise: ghost
 * exchange (two small kernels, or ncclSend/ncclRecv) on a private stream overlapped with the interior rows, then the
 * <= kl+ku boundary rows.  No all-reduce/all-gather on the path.
 * Shards: rows are split in blocks of per = roundup4(ceil(n/world)); trailing ranks may own NOTHING (they have no
 * neighbours and every sharded call returns 0 on them) and the last non-empty shard may be short.  A partition whose
 * full shards are narrower than the halo (per < max(kl,ku) with more than one non-empty rank) is rejected with -2 on
 * EVERY rank alike.  A device-side ghost wait that times out (lbm_set_tuning "halo_timeout_ms", default 30000) never
 * traps: it sets a fault word that the next sharded call / lbm_sync returns as 2001. */
int lbm_shard_plan(int64_t n, int64_t kl, int64_t ku, int rank, int world, int64_t *out8);
int lbm_comm_unique_id(void *id128);
int lbm_comm_init(lbm_handle_t h, const void *id128, int rank, int world);
int lbm_comm_destroy(lbm_handle_t h);
/* Peer-memory ghost transport (optional, one box): after lbm_comm_init every rank calls lbm_peer_alloc (a mailbox
 * block of 4 slots of slot_bytes each in its own HBM; returns the 64-byte CUDA IPC handle), the host exchanges the
 * handles (all-gather over the process group), and every rank calls lbm_peer_connect with its neighbours' handles
 * (NULL at the ends).  From then on the halo exchange of every *_sharded call and of lbm_halo_exchange is two tiny
 * kernels doing NVLink peer stores + flag handshakes instead of an NCCL send/recv group; halos larger than
 * slot_bytes, or lbm_set_tuning("exchange_mode", 1), fall back to NCCL.  lbm_peer_active: 1 if the peer path is on. */
int lbm_peer_alloc(lbm_handle_t h, size_t slot_bytes, void *ipc_handle64);
int lbm_peer_connect(lbm_handle_t h, const void *left_handle64, const void *right_handle64);
int lbm_peer_disconnect(lbm_handle_t h);
int lbm_peer_active(lbm_handle_t h);
int lbm_halo_exchange(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, void *xbuf, int elem_bytes);
int lbm_dgbmv_sharded(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, double alpha, const double *A_slab,
                      int64_t lda, double *xbuf, double beta, double *y_own, int overlap);
int lbm_sgbmv_sharded(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, float alpha, const float *A_slab,
                      int64_t lda, float *xbuf, float beta, float *y_own, int overlap);

/* mul!(y, ApplyMatrix(*, F_0, ..., F_{nf-1}), x) row-sharded in ONE call: right to left, the output of stage f is written
 * into the owned block of a ghost-cell vector laid out for the factor that consumes it, whose apply exchanges that
 * intermediate's halo (north_star: "halos of the intermediate in product chains"; flattening rule of
 * /root/reference/src/blockkron.jl:208-214).  Factor f is n x n with bandwidths (kl[f],ku[f]); A_slab[f] is this rank's
 * column slice of its band array (lbm_shard_plan(n,kl[f],ku[f],...)), xbuf the ghost-cell vector of factor nf-1, y_own
 * the owned rows.  Temporaries live in the handle's scratch. */
int lbm_dgbmv_chain_sharded(lbm_handle_t h, int nf, int64_t n, const int64_t *kl, const int64_t *ku,
                            const double *const *A_slab, const int64_t *lda, double alpha, double *xbuf, double beta,
                            double *y_own);
int lbm_sgbmv_chain_sharded(lbm_handle_t h, int nf, int64_t n, const int64_t *kl, const int64_t *ku,
                            const float *const *A_slab, const int64_t *lda, float alpha, float *xbuf, float beta,
                            float *y_own);
/* Tridiagonal / SymTridiagonal / Bidiagonal mat-vec row-sharded in ONE call (BASELINE config 2): plan =
 * lbm_shard_plan(n, dl_ext != NULL, du_ext != NULL, ...); d_ext = d[c0:c1], dl_ext = dl[c0:c1-1], du_ext = du[c0:c1-1]
 * (the diagonals over the ghost-extended index range; SymTridiagonal passes dl_ext == du_ext); xbuf = ghost-cell vector
 * of length c1-c0; y_own = rows [r0,r1).  overlap = 1: the one-element ghost exchange is fused into the launch. */
int lbm_dtridiag_mv_sharded(lbm_handle_t h, int64_t n, const double *dl_ext, const double *d_ext, const double *du_ext,
                            double alpha, double *xbuf, double beta, double *y_own, int overlap);
int lbm_stridiag_mv_sharded(lbm_handle_t h, int64_t n, const float *dl_ext, const float *d_ext, const float *du_ext,
                            float alpha, float *xbuf, float beta, float *y_own, int overlap);

/* kron(A,I)+kron(I,B) partitioned by grid columns: plan = lbm_shard_plan(n1, kla, kua, ...) over COLUMNS
 * of X; A_slab = columns [c0,c1) of A's band array; xbuf = n2 x (c1-c0) ghost-cell buffer
 * [left ghost columns | owned | right ghost columns]; y_own = n2 x (r1-r0).  Ghost columns travel by
 * ncclSend/ncclRecv while the interior columns are computed; B*X is rank-local. */
int lbm_dkronsum2d_mv_sharded(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua,
                              const double *A_slab, int64_t lda, int64_t klb, int64_t kub, const double *B,
                              int64_t ldb, double alpha, double *xbuf, double beta, double *y_own, int overlap);
int lbm_skronsum2d_mv_sharded(lbm_handle_t h, int64_t n1, int64_t n2, int64_t kla, int64_t kua,
                              const float *A_slab, int64_t lda, int64_t klb, int64_t kub, const float *B,
                              int64_t ldb, float alpha, float *xbuf, float beta, float *y_own, int overlap);

/* ---- single-process multi-device mode ("one host thread driving every GPU") -------------------------- */
/* For a host that is ONE process (a Julia session): no MPI.jl / Distributed.jl, no rendezvous, no NCCL.
 * lbm_create_multi creates one handle + stream per device, enables peer access between row-neighbours and cross-links
 * their mailbox blocks by plain peer pointers; every call below ENQUEUES on all devices from the calling thread and
 * returns (lbm_multi_sync waits).  The per-device handles (lbm_multi_handle) carry rank / world, so every `*_sharded`
 * entry point above also works on them (buffers from lbm_malloc on that handle) -- e.g. lbm_?tridiag_mv_sharded and
 * lbm_?kronsum2d_mv_sharded; call them device by device (for a multi-stage operation: all devices of one stage before
 * the next stage).  This is the object behind a sharded device layout in the Julia shim: `materialize!(::MulAdd{...})`
 * (hooks of /root/reference/src/tridiag.jl:685-688, src/bidiag.jl:438-440) forwards to lbm_?gbmv_sharded_all. */
typedef struct lbm_multi_s *lbm_multi_t;
typedef struct lbm_shard_s *lbm_shard_t;   /* row partition of an n x n banded operator + its per-device slabs */
typedef struct lbm_svec_s *lbm_svec_t;     /* a length-n vector in the ghost-cell layout of a shard */
int lbm_create_multi(lbm_multi_t *out, int ndev, const int *devices);   /* devices == NULL: 0 .. ndev-1 */
int lbm_destroy_multi(lbm_multi_t m);
int lbm_multi_ndev(lbm_multi_t m);
lbm_handle_t lbm_multi_handle(lbm_multi_t m, int r);
int lbm_multi_sync(lbm_multi_t m);                       /* waits for every device; reports device-side faults */
const char *lbm_multi_last_error_string(lbm_multi_t m);
/* elem_bytes: 8 = Float64, 4 = Float32.  The shard OWNS its device memory (the library allocates the slabs). */
int lbm_shard_create(lbm_multi_t m, int elem_bytes, int64_t n, int64_t kl, int64_t ku, lbm_shard_t *out);
int lbm_shard_destroy(lbm_shard_t s);
/* A_host: the (kl+ku+1) x n diagonal-major array as BandedMatrices stores it (leading dimension lda_host); device r
 * receives the column slice [c0,c1) of lbm_shard_plan. */
int lbm_shard_distribute(lbm_shard_t s, const void *A_host, int64_t lda_host);
int lbm_shard_fill_uniform(lbm_shard_t s, uint64_t seed);          /* brand(n,n,kl,ku) from the counter-based stream */
int lbm_shard_slab(lbm_shard_t s, int r, void **dptr, int64_t *plan8);
int lbm_svec_create(lbm_shard_t s, lbm_svec_t *out);
int lbm_svec_destroy(lbm_svec_t v);
int lbm_svec_distribute(lbm_svec_t v, const void *x_host);         /* n host elements -> owned blocks */
int lbm_svec_gather(lbm_svec_t v, void *y_host);                   /* owned blocks -> n host elements */
int lbm_svec_fill_uniform(lbm_svec_t v, uint64_t seed);
int lbm_svec_buffer(lbm_svec_t v, int r, void **ghost_buf, void **owned);
/* y <- alpha*A*x + beta*y on all devices: one launch per device, ghost exchange fused into it.  x must be a vector of A's
 * own layout (lbm_svec_create(A, ...)); y any vector of the same n; -3 / -5 = DimensionMismatch. */
int lbm_dgbmv_sharded_all(lbm_shard_t A, double alpha, lbm_svec_t x, double beta, lbm_svec_t y);
int lbm_sgbmv_sharded_all(lbm_shard_t A, float alpha, lbm_svec_t x, float beta, lbm_svec_t y);
/* y <- alpha*F_0*(F_1*(...F_{nf-1}*x)) + beta*y, every factor sharded over the same devices; x in F_{nf-1}'s layout;
 * the halos of each intermediate are exchanged by the stage that consumes it. */
int lbm_dgbmv_chain_sharded_all(int nf, const lbm_shard_t *F, double alpha, lbm_svec_t x, double beta, lbm_svec_t y);
int lbm_sgbmv_chain_sharded_all(int nf, const lbm_shard_t *F, float alpha, lbm_svec_t x, float beta, lbm_svec_t y);

/* ---- structure helpers (pure integer, host side, no GPU needed) ------------------------ */
/* bandwidths(ApplyMatrix(*, F1..Fnf)) = min.(size-1, (sum kl, sum ku))  ([up] LazyArrays;
 * reference README.md:13-26 shows (1,1)*(1,1) -> (2,2)). */
int lbm_prod_bandwidths(int64_t m, int64_t n, int nf, const int64_t *kl, const int64_t *ku,
                        int64_t *out_kl, int64_t *out_ku);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
