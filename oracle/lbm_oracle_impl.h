/*
 * lbm_oracle_impl.h -- type-generic body of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Included twice by lbm_oracle.c with
 *     T   = double / float
 *     FN  = name-mangling macro (oracle_d## / oracle_s##)
 *
 * Every routine restates, in plain sequential-per-row C, what the reference's
 * path computes.  The arithmetic of that path lives in un-vendored Julia
 * dependencies (BandedMatrices.jl >= 1.0 `gbmv!`/`gbmm!`, ArrayLayouts.jl >= 1.2.1
 * `MulAdd`, LazyArrays.jl >= 2.8 `Mul` flattening, BlockArrays.jl >= 1.0
 * `BlockKron`, pinned only by /root/reference/Project.toml:24-34) and in
 * OpenBLAS `?gbmv_`; what IS in the reference are the element definitions and
 * the Kronecker identity, which are cited per function below.
 *
 * Band storage (BandedMatrices / LAPACK "AB" format):
 *     A[i,j] = data[(u + i - j) + lda*j]   (0-based),   -u <= i-j <= l
 * Out-of-matrix slots of `data` are NEVER read (they may hold NaN garbage).
 */

/* y <- alpha*op(A)*x + beta*y for an m x n banded A, op = N / T.
 * Semantics of BLAS xGBMV, which BandedMatrices.gbmv! ccalls;
 * beta == 0 means y is not read (NaNs in y must not propagate). */
int FN(gbmv)(char trans, int64_t m, int64_t n, int64_t kl, int64_t ku,
             T alpha, const T *A, int64_t lda, const T *x, T beta, T *y)
{
    if (m < 0 || n < 0 || kl < 0 || ku < 0 || lda < kl + ku + 1) return -1;
    const int tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!tr) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < m; ++i) {
            int64_t jlo = i - kl < 0 ? 0 : i - kl;
            int64_t jhi = i + ku > n - 1 ? n - 1 : i + ku;
            T acc = 0;
            for (int64_t j = jlo; j <= jhi; ++j)
                acc += A[(ku + i - j) + lda * j] * x[j];
            y[i] = (beta == (T)0) ? alpha * acc : alpha * acc + beta * y[i];
        }
    } else {
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < n; ++j) {
            int64_t ilo = j - ku < 0 ? 0 : j - ku;
            int64_t ihi = j + kl > m - 1 ? m - 1 : j + kl;
            T acc = 0;
            for (int64_t i = ilo; i <= ihi; ++i)
                acc += A[(ku + i - j) + lda * j] * x[i];
            y[j] = (beta == (T)0) ? alpha * acc : alpha * acc + beta * y[j];
        }
    }
    return 0;
}

/* Y <- alpha*A*X + beta*Y for A = Tridiagonal(dl,d,du) (n x n), X n x nrhs.
 * Element definitions: /root/reference/src/tridiag.jl:471-484 (Tridiagonal:
 * A[i,i]=d[i], A[i+1,i]=dl[i], A[i,i+1]=du[i]) and :301-314 (SymTridiagonal:
 * dl == du == ev, only ev[1:n-1] used, src/tridiag.jl:11,133-137);
 * Bidiagonal (src/bidiag.jl:111-124): uplo 'U' -> dl == NULL, 'L' -> du == NULL.
 * Same three-point form as the in-repo stencil loop src/tridiag.jl:665-682. */
int FN(tridiag_mv)(int64_t n, const T *dl, const T *d, const T *du, T alpha,
                   const T *X, int64_t ldx, int64_t nrhs, T beta, T *Y, int64_t ldy)
{
    if (n < 0 || nrhs < 0) return -1;
    for (int64_t c = 0; c < nrhs; ++c) {
        const T *x = X + c * ldx;
        T *y = Y + c * ldy;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            T acc = 0;
            if (dl && i > 0) acc += dl[i - 1] * x[i - 1];
            acc += d[i] * x[i];
            if (du && i < n - 1) acc += du[i] * x[i + 1];
            y[i] = (beta == (T)0) ? alpha * acc : alpha * acc + beta * y[i];
        }
    }
    return 0;
}

/* C <- alpha*A*B + beta*C, A m x k (kla,kua), B k x n (klb,kub), C m x n stored
 * with band (klc,kuc).  C[i,j] = sum_{p} A[i,p] B[p,j],
 * p in [max(i-kla, j-kub, 0), min(i+kua, j+klb, k-1)] -- the banded product
 * behind `ApplyMatrix(*, A, B)` materialisation (reference README.md:8-26;
 * BandedMatrices gbmm! issues one gbmv per column of B, i.e.
 * for fixed j it sums over p in increasing order, as here).
 * In-matrix slots of C's band outside the product band get beta*C (+0);
 * out-of-matrix slots of C are set to 0 when beta == 0 and left alone otherwise. */
int FN(gbmm)(int64_t m, int64_t n, int64_t k, T alpha,
             int64_t kla, int64_t kua, const T *A, int64_t lda,
             int64_t klb, int64_t kub, const T *B, int64_t ldb, T beta,
             int64_t klc, int64_t kuc, T *C, int64_t ldc)
{
    if (m < 0 || n < 0 || k < 0) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) {
        for (int64_t r = 0; r < klc + kuc + 1; ++r) {
            int64_t i = j + r - kuc;
            T *c = &C[r + ldc * j];
            if (i < 0 || i >= m) {
                if (beta == (T)0) *c = 0;
                continue;
            }
            int64_t plo = i - kla > j - kub ? i - kla : j - kub;
            int64_t phi = i + kua < j + klb ? i + kua : j + klb;
            if (plo < 0) plo = 0;
            if (phi > k - 1) phi = k - 1;
            T acc = 0;
            for (int64_t p = plo; p <= phi; ++p)
                acc += A[(kua + i - p) + lda * p] * B[(kub + p - j) + ldb * j];
            *c = (beta == (T)0) ? alpha * acc : alpha * acc + beta * (*c);
        }
    }
    return 0;
}

/* Batched three-factor chain D_b = (A_b*B_b)*C_b, each factor n x n with tight
 * lda = kl+ku+1; pairwise left-to-right with a full intermediate, as the
 * reference's path materialises chains. */
int FN(gbmm_chain3_batched)(int64_t batch, int64_t n, const int64_t *kl, const int64_t *ku,
                            const T *A, int64_t sA, const T *B, int64_t sB,
                            const T *C, int64_t sC, T *D, int64_t sD)
{
    int64_t l01 = kl[0] + kl[1], u01 = ku[0] + ku[1];
    int64_t l3 = l01 + kl[2], u3 = u01 + ku[2];
    int64_t ldt = l01 + u01 + 1;
    T *tmp = (T *)malloc(sizeof(T) * (size_t)(ldt * (n > 0 ? n : 1)));
    if (!tmp) return 1;
    for (int64_t b = 0; b < batch; ++b) {
        FN(gbmm)(n, n, n, (T)1, kl[0], ku[0], A + b * sA, kl[0] + ku[0] + 1,
                 kl[1], ku[1], B + b * sB, kl[1] + ku[1] + 1, (T)0, l01, u01, tmp, ldt);
        FN(gbmm)(n, n, n, (T)1, l01, u01, tmp, ldt,
                 kl[2], ku[2], C + b * sC, kl[2] + ku[2] + 1, (T)0, l3, u3, D + b * sD, l3 + u3 + 1);
    }
    free(tmp);
    return 0;
}

/* y <- alpha*(kron(A, I_n2) + kron(I_n1, B))*x + beta*y with x = vec(X),
 * X n2 x n1 column-major: result vec(X*A^T + B*X).
 * Kronecker identity used by the reference: /root/reference/src/blockkron.jl:272-278
 * and :440 (`KronTrav(A,B)*DiagTrav(X) = DiagTrav(B*X*A')`); BlockKron(A,B) has
 * blocks A[K,J]*B (test/test_blockkron.jl:94-107), so BlockKron(A,I)+BlockKron(I,B)
 * is exactly this operator (test/test_blockkron.jl:296-306, the 2-D Laplacian).
 * A is n1 x n1 banded (kla,kua), B is n2 x n2 banded (klb,kub). */
int FN(kronsum2d_mv)(int64_t n1, int64_t n2,
                     int64_t kla, int64_t kua, const T *A, int64_t lda,
                     int64_t klb, int64_t kub, const T *B, int64_t ldb,
                     T alpha, const T *x, T beta, T *y)
{
    if (n1 < 0 || n2 < 0) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < n1; ++c) {
        for (int64_t r = 0; r < n2; ++r) {
            T acc = 0;
            /* (X A^T)[r,c] = sum_q X[r,q] A[c,q] */
            int64_t qlo = c - kla < 0 ? 0 : c - kla;
            int64_t qhi = c + kua > n1 - 1 ? n1 - 1 : c + kua;
            for (int64_t q = qlo; q <= qhi; ++q)
                acc += A[(kua + c - q) + lda * q] * x[r + n2 * q];
            /* (B X)[r,c] = sum_p B[r,p] X[p,c] */
            int64_t plo = r - klb < 0 ? 0 : r - klb;
            int64_t phi = r + kub > n2 - 1 ? n2 - 1 : r + kub;
            for (int64_t p = plo; p <= phi; ++p)
                acc += B[(kub + r - p) + ldb * p] * x[p + n2 * c];
            int64_t o = r + n2 * c;
            y[o] = (beta == (T)0) ? alpha * acc : alpha * acc + beta * y[o];
        }
    }
    return 0;
}

/* y <- alpha*kron(A,B)*x + beta*y = alpha*vec(B*X*A^T) + beta*y, A n1 x n1, B n2 x n2
 * banded; Z = B*X first, then Y = Z*A^T (src/blockkron.jl:440 evaluates B*X*A'
 * left to right).  Never forms the Kronecker product. */
int FN(kron2d_mv)(int64_t n1, int64_t n2,
                  int64_t kla, int64_t kua, const T *A, int64_t lda,
                  int64_t klb, int64_t kub, const T *B, int64_t ldb,
                  T alpha, const T *x, T beta, T *y)
{
    if (n1 < 0 || n2 < 0) return -1;
    T *Z = (T *)malloc(sizeof(T) * (size_t)(n1 * n2 > 0 ? n1 * n2 : 1));
    if (!Z) return 1;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < n1; ++c)
        for (int64_t r = 0; r < n2; ++r) {
            int64_t plo = r - klb < 0 ? 0 : r - klb;
            int64_t phi = r + kub > n2 - 1 ? n2 - 1 : r + kub;
            T acc = 0;
            for (int64_t p = plo; p <= phi; ++p)
                acc += B[(kub + r - p) + ldb * p] * x[p + n2 * c];
            Z[r + n2 * c] = acc;
        }
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < n1; ++c)
        for (int64_t r = 0; r < n2; ++r) {
            int64_t qlo = c - kla < 0 ? 0 : c - kla;
            int64_t qhi = c + kua > n1 - 1 ? n1 - 1 : c + kua;
            T acc = 0;
            for (int64_t q = qlo; q <= qhi; ++q)
                acc += Z[r + n2 * q] * A[(kua + c - q) + lda * q];
            int64_t o = r + n2 * c;
            y[o] = (beta == (T)0) ? alpha * acc : alpha * acc + beta * y[o];
        }
    free(Z);
    return 0;
}

/* dot(x, A, y) for A = Tridiagonal(dl,d,du): the sliding-window loop of
 * /root/reference/src/tridiag.jl:665-682, i.e. sum_j (A' x)[j] * y[j]. */
T FN(tridiag_dot)(int64_t n, const T *x, const T *dl, const T *d, const T *du, const T *y)
{
    if (n <= 0) return 0;
    if (n == 1) return x[0] * d[0] * y[0];
    T r = 0;
    for (int64_t j = 0; j < n; ++j) {
        T s = d[j] * x[j];
        if (j > 0) s += du[j - 1] * x[j - 1];
        if (j < n - 1) s += dl[j] * x[j + 1];
        r += s * y[j];
    }
    return r;
}
