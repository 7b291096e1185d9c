/*
 * lbm_oracle.c -- CPU ORACLE for the banded mul!/materialize hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The
 * product (lazybandedmatrices.jl_b200/) never links, imports or calls it and has
 * no CPU fallback.
 *
 * PARITY STATUS.  The reference (LazyBandedMatrices.jl v0.11.9, pure Julia)
 * cannot be executed here (no julia in the image) and the arithmetic of its
 * banded `mul!`/`materialize` lives in un-vendored dependencies that are absent
 * from /root/reference (Project.toml:6-16, compat-pinned only at :24-34):
 *   BandedMatrices >= 1.0 (<2)  gbmv!/gbmm!     ->  OpenBLAS ?gbmv_ (libblastrampoline)
 *   ArrayLayouts  >= 1.2.1      Mul/MulAdd/materialize!
 *   LazyArrays    >= 2.8 (<3)   ApplyArray, Mul flattening
 *   BlockArrays   >= 1.0, BlockBandedMatrices 0.13   BlockKron, blockbandwidths
 * This file restates their published algorithms.  It is pinned by
 *   (1) every golden vector / known-answer test the reference's own tests hold
 *       for the path (tests/golden/reference_kats.json, transcribed with
 *       file:line from /root/reference/test/ (the .jl test files); checked in
 *       tests/test_oracle_golden.py), and
 *   (2) OpenBLAS dgbmv/sgbmv through scipy.linalg.blas -- the very BLAS routine
 *       the reference's gbmv! ccalls (tests/test_oracle_blas.py).
 * The VALUES of ApplyMatrix(*, brand, brand) and of BandedMatrix gbmv on random
 * data are NOT pinned by any reference test (the reference's random inputs come
 * from Julia's RNG): for those rows the header says it plainly -- "parity
 * unpinned" by reference vectors; pinned by (2), by dense numpy matmul on small
 * sizes, and by the mathematical definition only.
 *
 * Build: make -C oracle   ->  oracle/_build/liblbm_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- counter-based generator ------------------------------
 * v(idx) = the (idx+1)-th output of a splitmix64 stream seeded with `seed`,
 * mapped to [0,1): top 53 bits * 2^-53 (double) or top 24 bits * 2^-24 (float).
 * Random access, so CPU and GPU generate identical data without a transfer. */
static inline uint64_t lbm_mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t lbm_sm64(uint64_t seed, uint64_t idx)
{
    return lbm_mix64(seed + (idx + 1ULL) * 0x9E3779B97F4A7C15ULL);
}

int oracle_dfill_uniform(uint64_t seed, int64_t offset, int64_t count, double *dst)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < count; ++i)
        dst[i] = (double)(lbm_sm64(seed, (uint64_t)(offset + i)) >> 11) * 0x1.0p-53;
    return 0;
}
int oracle_sfill_uniform(uint64_t seed, int64_t offset, int64_t count, float *dst)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < count; ++i)
        dst[i] = (float)(lbm_sm64(seed, (uint64_t)(offset + i)) >> 40) * 0x1.0p-24f;
    return 0;
}

#define T double
#define FN(name) oracle_d##name
#include "lbm_oracle_impl.h"
#undef T
#undef FN

#define T float
#define FN(name) oracle_s##name
#include "lbm_oracle_impl.h"
#undef T
#undef FN

int oracle_num_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
