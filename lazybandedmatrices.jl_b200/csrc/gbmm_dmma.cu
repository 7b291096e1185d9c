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
// shared memory through a 3-stage cp.async pipeline (16-byte copies, 8-byte zero-filling copies
// at the band edges do the masking; pitches 72 / 20 doubles make both the staging writes and the DMMA fragment loads
// bank-conflict-free), and 4 warps each
// keep a 32x32 block of C (16 DMMA accumulators) in registers.  8x8x4 sub-products whose A or B
// block lies wholly outside its band are skipped (warp-uniform test), which removes most of
// the padding waste of the parallelogram.
#include "lbm_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, KC = 16, NT = 128;
constexpr int PA = 68;  // sA[p][i], i fastest; 2*PA = 8 (mod 32): each half-warp's 16 fragment loads hit 16 bank pairs
constexpr int PB = 20;  // sB[j][p], p fastest; 2*PB = 8 (mod 32), same property

struct DmmaArgs {
    int64_t m, n, k;
    int64_t kla, kua, lda, klb, kub, ldb, klc, kuc, ldc;
    double alpha, beta;
    const double *A;
    const double *B;
    double *C;
    int64_t gy;  // row tiles per column block
    int flags;   // 1: rotate the warp -> quadrant map per CTA, 2: interleave the two halves of the p-range
    int tpc;     // warp-specialised kernel: consecutive row tiles per CTA
    // corner launch (MODE 2) of the plain kernel: it covers the column blocks [0, cb_lo) and [cb_hi0, gx) only -- every tile
    // of the blocks in between is an interior tile the warp-specialised kernel has already done
    int64_t cb_lo, cb_hi0;
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

constexpr int STAGE_DBL = KC * PA + TN * PB;  // doubles per stage

// Warp -> block map inside the 64x64 tile: warp (a,b) owns the 8-row blocks a0 + BSX*x and the 8-column blocks
// b0 + BSX*y (x, y < 4).  BSX = 1: quadrants.  BSX = 2 interleaves the warps' blocks so that all four warps see the
// same band-edge pattern in the ramp chunks (busiest warp +6 % over the mean instead of +26 %), but fewer chunks then
// take the unmasked path; measured 18.5 vs 19.9 TFLOP/s, so quadrants stay (profiles/dmma_r01_notes.md).
constexpr int BSX = 1;
// Block-validity masks of a warp panel for one chunk.  Bit 4*ks + x stands for the 8x4 (A) / 4x8 (B) DMMA operand
// block x of k-step ks; with t = 2*BSX*x - ks in [-3, 6*BSX] the block touches the band iff tlo <= t <= thi, where
// tlo / thi are two integers per chunk (see the kernel).  MASK_GE[c + 3] = {t >= c}, c in [-3, 6*BSX + 1];
// MASK_LE[c + 4] = {t <= c}, c in [-4, 6*BSX].
constexpr int TMAX = 6 * BSX;
constexpr unsigned mask_cmp(int c, bool ge) {
    unsigned m = 0;
    for (int ks = 0; ks < 4; ++ks)
        for (int x = 0; x < 4; ++x) {
            const int t = 2 * BSX * x - ks;
            if (ge ? (t >= c) : (t <= c)) m |= 1u << (4 * ks + x);
        }
    return m;
}
struct MaskTables {
    unsigned short ge[TMAX + 5], le[TMAX + 5];
};
constexpr MaskTables make_tables() {
    MaskTables t{};
    for (int c = -3; c <= TMAX + 1; ++c) t.ge[c + 3] = (unsigned short)mask_cmp(c, true);
    for (int c = -4; c <= TMAX; ++c) t.le[c + 4] = (unsigned short)mask_cmp(c, false);
    return t;
}
__constant__ MaskTables MASKS = make_tables();
__device__ __forceinline__ int fdiv4(int a) { return a >> 2; }             // floor(a / 4)
__device__ __forceinline__ int cdiv4(int a) { return (a + 3) >> 2; }       // ceil(a / 4)
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// blocks with  lo4 <= 4t <= hi4
__device__ __forceinline__ unsigned band_mask(int lo4, int hi4) {
    return (unsigned)MASKS.ge[clampi(cdiv4(lo4), -3, TMAX + 1) + 3] & (unsigned)MASKS.le[clampi(fdiv4(hi4), -4, TMAX) + 4];
}

// MODE 0: every tile (one launch).  MODE 1: interior tiles only -- the lean instantiation (no zero-filling staging,
// no per-block tests) that carries ~all of the work; MODE 2: the remaining corner tiles.
template <int NSTAGE, int MINB, int MODE>
__global__ void __launch_bounds__(NT, MINB) gbmm_dmma_kernel(const DmmaArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // tile order: the tiles of one column block are consecutive CTAs (they share the B tile in
    // L2) and consecutive column blocks share most of their A rows
    const int64_t lin = (int64_t)blockIdx.x + (int64_t)gridDim.x * blockIdx.y;
    // warp's 32x32 quadrant of the tile; rotated per CTA so that the quadrant that is busiest in the ramp chunks
    // does not always land on the same SM sub-core (warp w of every resident CTA shares sub-core w % 4)
    const int role = (a.flags & 1) ? ((warp + (int)(lin & 3)) & 3) : warp;
    constexpr int WOFF = BSX == 1 ? 32 : 8, BST = 8 * BSX, SPAN = 24 * BSX + 7;  // first block, block stride, panel span - 1
    const int wi = (role & 1) * WOFF, wj = (role >> 1) * WOFF;
    const int64_t by = lin % a.gy;
    int64_t cb = lin / a.gy;
    if (MODE == 2 && cb >= a.cb_lo) cb = a.cb_hi0 + (cb - a.cb_lo);
    const int64_t j0 = cb * TN;
    if (j0 >= a.n) return;
    int64_t I0 = floordiv(j0 - a.kuc, TM);
    if (I0 < 0) I0 = 0;
    const int64_t i0 = (I0 + by) * TM;
    if (i0 >= a.m || i0 > j0 + TN - 1 + a.klc) return;

    // p-range of the tile
    int64_t plo = i0 - a.kla > j0 - a.kub ? i0 - a.kla : j0 - a.kub;
    int64_t phi = i0 + TM - 1 + a.kua < j0 + TN - 1 + a.klb ? i0 + TM - 1 + a.kua : j0 + TN - 1 + a.klb;
    if (plo < 0) plo = 0;
    if (phi > a.k - 1) phi = a.k - 1;
    plo &= ~(int64_t)3;
    const int nch = phi >= plo ? (int)((phi - plo + KC) / KC) : 0;
    // Interior tiles (all but the matrix corners): every dense position (i,p) / (p,j) the chunks touch -- in band or
    // not -- is an address inside the arrays, so the tiles are staged RAW with unpredicated 16-byte copies and the
    // band edges are masked in registers (partial chunks only).  Corner tiles keep the zero-filling staging.
    const bool interior_t = i0 + TM <= a.m && j0 >= 4 && j0 + TN + 4 <= a.n && plo >= 4 && plo + (int64_t)nch * KC + 4 <= a.k &&
                          a.lda >= 33 && a.ldb >= 33 && ((a.lda - 1) & 1) == 0 && ((a.ldb - 1) & 1) == 0 &&
                          ((reinterpret_cast<uintptr_t>(a.A + a.kua + i0 + (a.lda - 1) * plo) & 15u) == 0) &&
                          ((reinterpret_cast<uintptr_t>(a.B + a.kub + plo + (a.ldb - 1) * j0) & 15u) == 0);
    if (MODE == 1 && !interior_t) return;
    if (MODE == 2 && interior_t) return;
    const bool interior = MODE == 1 ? true : (MODE == 2 ? false : interior_t);

    double acc[4][4][2];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;

    // ---- staging maps (all offsets relative to the tile: 32-bit) ---------------------------
    // Elements are staged in PAIRS along the direction that is contiguous in global memory:
    // one 16-byte cp.async when both are in-band and the source is 16-byte aligned (always the
    // case for odd lda with even ku, e.g. l=u=128), else two 8-byte zero-filling copies.
    // A: rows (2*ii2, 2*ii2+1), ii2 = tid % 32; p = tid / 32 + 4q (q < 4).
    //    A[i,p] = Ad[(kua + i) + (lda-1)*p], valid for p in [i - kla, i + kua], p < k, i < m.
    // B: p pair (2*pp2, 2*pp2+1), pp2 = tid % 8; j = tid / 8 + 16q (q < 4).
    //    B[p,j] = Bd[(kub + p) + (ldb-1)*j], valid for p in [j - kub, j + klb], p < k, j < n.
    const int a_ii = 2 * (tid & 31), a_p0 = tid >> 5;
    const int b_pp = 2 * (tid & 7), b_j0 = tid >> 3;
    const int kmax = (int)(a.k - 1 - plo);             // relative p must be <= kmax
    const bool a_ok0 = (i0 + a_ii) < a.m, a_ok1 = (i0 + a_ii + 1) < a.m;
    const int a_lo0 = (int)(i0 + a_ii - a.kla - plo);  // valid relative p range of row a_ii (row a_ii+1: +1)
    const int a_hi0 = (int)(i0 + a_ii + a.kua - plo);
    const double *a_ptr = a.A + (a.kua + i0 + a_ii) + (a.lda - 1) * plo;  // + (lda-1)*prel
    const int64_t a_step = a.lda - 1;
    const int b_lo0 = (int)(j0 + b_j0 - a.kub - plo);  // valid relative p range of column b_j0 (+16q for q)
    const int b_hi0 = (int)(j0 + b_j0 + a.klb - plo);
    const int b_ncol = (int)((a.n - j0 - b_j0 + 15) / 16);  // q < b_ncol  <=>  column j0 + b_j0 + 16q < n
    const double *b_ptr = a.B + (a.kub + plo + b_pp) + (a.ldb - 1) * (j0 + b_j0);  // + prel + (ldb-1)*16q
    const int64_t b_step = (a.ldb - 1) * 16;

    auto cp16 = [&](double *dst, const double *src) {
        const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    };
    auto stage_pair = [&](double *dst, const double *src, bool ok0, bool ok1, const double *safe) {
        if (ok0 && ok1 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
            cp16(dst, src);
        } else {
            cp_async8_zfill(dst, ok0 ? src : safe, ok0);
            cp_async8_zfill(dst + 1, ok1 ? src + 1 : safe, ok1);
        }
    };
    // 16-byte alignment of every pair source (true for l=u=128: lda-1, ldb-1 and ku even)
    const bool a_al = ((reinterpret_cast<uintptr_t>(a_ptr) & 15u) == 0) && ((a_step & 1) == 0);
    const bool b_al = ((reinterpret_cast<uintptr_t>(b_ptr) & 15u) == 0) && (((a.ldb - 1) & 1) == 0);
    // tile-level band limits (relative p) for the "whole chunk in band" fast paths
    const int ti0 = (int)(i0 - plo), tj0 = (int)(j0 - plo);
    const bool rows_in = i0 + TM <= a.m, cols_in = j0 + TN <= a.n;
    const int kla = (int)a.kla, kua = (int)a.kua, klb = (int)a.klb, kub = (int)a.kub;

    // chunk visiting order: it -> chunk.  Interleaving the two halves of the p-range pairs every cheap ramp chunk
    // (few blocks in band) with an expensive mid-range chunk, which keeps the prefetch distance useful.
    const int half = (nch + 1) >> 1;
    const bool reorder = (a.flags & 2) != 0;
    auto chunk_of = [&](int it) { return reorder ? ((it & 1) ? half + (it >> 1) : (it >> 1)) : it; };
    auto issue = [&](int it) {
        const int ch = chunk_of(it);
        double *sA = smem + (it % NSTAGE) * STAGE_DBL;
        double *sB = sA + KC * PA;
        const int pr = ch * KC;  // relative p of the chunk start
        if (interior) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cp16(sA + (a_p0 + 4 * q) * PA + a_ii, a_ptr + a_step * (pr + a_p0 + 4 * q));
#pragma unroll
            for (int q = 0; q < 4; ++q) cp16(sB + (b_j0 + 16 * q) * PB + b_pp, b_ptr + pr + b_step * q);
            return;
        }
        const bool p_in = pr + KC - 1 <= kmax;
        // every (i,p) of the 64 x 16 A tile in band  <=>  ti0+63 - pr <= kla  and  ti0 - (pr+15) >= -kua
        if (a_al && rows_in && p_in && ti0 + TM - 1 - pr <= kla && ti0 - (pr + KC - 1) >= -kua) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cp16(sA + (a_p0 + 4 * q) * PA + a_ii, a_ptr + a_step * (pr + a_p0 + 4 * q));
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int p = pr + a_p0 + 4 * q;
                const bool pk = p <= kmax;
                const bool ok0 = a_ok0 && pk && p >= a_lo0 && p <= a_hi0;
                const bool ok1 = a_ok1 && pk && p >= a_lo0 + 1 && p <= a_hi0 + 1;
                stage_pair(sA + (a_p0 + 4 * q) * PA + a_ii, a_ptr + a_step * p, ok0, ok1, a.A);
            }
        }
        // every (p,j) of the 16 x 64 B tile in band  <=>  (pr+15) - tj0 <= klb  and  pr - (tj0+63) >= -kub
        if (b_al && cols_in && p_in && (pr + KC - 1) - tj0 <= klb && pr - (tj0 + TN - 1) >= -kub) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cp16(sB + (b_j0 + 16 * q) * PB + b_pp, b_ptr + pr + b_step * q);
        } else {
            const int pb = pr + b_pp;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool cok = q < b_ncol;
                const int lo = b_lo0 + 16 * q, hi = b_hi0 + 16 * q;
                const bool ok0 = cok && pb >= lo && pb <= hi && pb <= kmax;
                const bool ok1 = cok && pb + 1 >= lo && pb + 1 <= hi && pb + 1 <= kmax;
                stage_pair(sB + (b_j0 + 16 * q) * PB + b_pp, b_ptr + pr + b_step * q, ok0, ok1, a.B);
            }
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
    const int fa_off = fc * PA + wi + fr;       // + 4*ks*PA + BST*x
    const int fb_off = (wj + fr) * PB + fc;     // + BST*x*PB + 4*ks

    for (int it = 0; it < nch; ++it) {
        cp_async_wait<NSTAGE - 2>();  // chunk `it` has landed (this thread's copies)
        __syncthreads();              // ... everyone's; and everyone is done computing chunk it-1
        if (it + NSTAGE - 1 < nch) issue(it + NSTAGE - 1);
        cp_async_commit();
        const double *sA = smem + (it % NSTAGE) * STAGE_DBL;
        const double *sB = sA + KC * PA;
        const int pr = chunk_of(it) * KC;
        constexpr bool DB = (MINB <= 3);  // double-buffer fragments only where registers allow it
        // whole 32 x 16 (A) and 16 x 32 (B) warp panels inside their bands: no per-block tests
        const bool full = (ia0 + SPAN - pr <= kla) && (ia0 - (pr + KC - 1) >= -kua) &&
                          ((pr + KC - 1) - jb0 <= klb) && (pr - (jb0 + SPAN) >= -kub);
        if (interior && !full) {
            // raw-staged partial chunk.  da = (first row of the warp panel) - (first p of the chunk); A block (x,ks)
            // holds i - p = da + BST*x - 4ks + [-3, 7]; db likewise for B with p - j = -(db + BST*y - 4ks) + [-7, 3]
            const int da = ia0 - pr, db = jb0 - pr;
            const unsigned mA = band_mask(-kua - 7 - da, kla + 3 - da);   // blocks touching A's band
            const unsigned mB = band_mask(-klb - 7 - db, kub + 3 - db);   // blocks touching B's band
            if (mA != 0 && mB != 0) {
                // per-lane element test: (unsigned)(d + kua) <= kla + kua  with d = i - p of the lane's element
                const int va = da + fr - fc + kua, wa = kla + kua;
                const int vb = -db + fc - fr + kub, wb = klb + kub;
#pragma unroll
                for (int ks = 0; ks < KC / 4; ++ks) {
                    const unsigned ma = (mA >> (4 * ks)) & 15u, mb = (mB >> (4 * ks)) & 15u;
                    if (ma == 0 || mb == 0) continue;
                    double ga[4], gb[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const double ra = sA[fa_off + 4 * ks * PA + BST * x];
                        const double rb = sB[fb_off + BST * x * PB + 4 * ks];
                        ga[x] = (unsigned)(va + BST * x - 4 * ks) <= (unsigned)wa ? ra : 0.0;
                        gb[x] = (unsigned)(vb - BST * x + 4 * ks) <= (unsigned)wb ? rb : 0.0;
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        if (!(ma >> x & 1u)) continue;
#pragma unroll
                        for (int y = 0; y < 4; ++y)
                            if (mb >> y & 1u) dmma884(acc[x][y][0], acc[x][y][1], ga[x], gb[y]);
                    }
                }
            }
            continue;
        }
        double fa[DB ? 2 : 1][4], fb[DB ? 2 : 1][4];
        if (DB) {
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                fa[0][x] = sA[fa_off + BST * x];
                fb[0][x] = sB[fb_off + BST * x * PB];
            }
        }
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            const int cur = DB ? (ks & 1) : 0, nxt = DB ? (cur ^ 1) : 0;
            if (DB) {
                if (ks + 1 < KC / 4) {  // fragments of the next k-step are in flight during this one's MMAs
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        fa[nxt][x] = sA[fa_off + 4 * (ks + 1) * PA + BST * x];
                        fb[nxt][x] = sB[fb_off + BST * x * PB + 4 * ks + 4];
                    }
                }
            } else {
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    fa[0][x] = sA[fa_off + 4 * ks * PA + BST * x];
                    fb[0][x] = sB[fb_off + BST * x * PB + 4 * ks];
                }
            }
            if (full) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], fa[cur][x], fb[cur][y]);
            } else {
                const int p = pr + 4 * ks;  // relative
                unsigned ma = 0, mb = 0;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int ia = ia0 + BST * x;
                    if (ia - (p + 3) <= kla && (ia + 7) - p >= -kua) ma |= 1u << x;
                    const int jb = jb0 + BST * x;
                    if ((p + 3) - jb >= -kub && p - (jb + 7) <= klb) mb |= 1u << x;
                }
                if (ma != 0 && mb != 0) {
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y)
                            if ((ma >> x & 1u) && (mb >> y & 1u))
                                dmma884(acc[x][y][0], acc[x][y][1], fa[cur][x], fb[cur][y]);
                }
            }
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
                const int64_t i = i0 + wi + BST * x + fr;
                const int64_t j = j0 + wj + BST * y + 2 * fc + e;
                const int64_t f = a.kuc + i - j;
                if (i < a.m && j < a.n && f >= 0 && f <= a.klc + a.kuc) {
                    double *c = a.C + f + a.ldc * j;
                    const double v = a.alpha * acc[x][y][e];
                    *c = (a.beta == 0.0) ? v : fma(a.beta, *c, v);
                }
            }
}

// ------------------------------------------------------------------------------------------
// column-block persistent variant: one CTA owns ALL row tiles of a 64-column block and runs the
// cp.async ring straight across tile boundaries (no pipeline fill/drain per 64x64 tile; the B
// column block is re-read from L1/L2 by the same CTA).
// ------------------------------------------------------------------------------------------
struct TileIter {            // walks (tile, chunk) of one column block; warp-uniform
    int64_t i0, pc, phi;     // current tile's first row, current chunk's first p, tile's last p
    int t, ntiles;
    bool valid;
};

template <int NSTAGE, int MINB>
__global__ void __launch_bounds__(NT, MINB) gbmm_dmma_col_kernel(const DmmaArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wi = (warp & 1) * 32, wj = (warp >> 1) * 32;
    const int64_t j0 = ((int64_t)blockIdx.x + (int64_t)gridDim.x * blockIdx.y) * TN;
    if (j0 >= a.n) return;
    int64_t I0 = floordiv(j0 - a.kuc, TM);
    if (I0 < 0) I0 = 0;
    // number of row tiles of this column block: i0 < m and i0 <= j0 + TN - 1 + klc
    int64_t last = j0 + TN - 1 + a.klc < a.m - 1 ? j0 + TN - 1 + a.klc : a.m - 1;
    const int ntiles = last >= I0 * TM ? (int)(last / TM - I0 + 1) : 0;
    const int kla = (int)a.kla, kua = (int)a.kua, klb = (int)a.klb, kub = (int)a.kub;

    auto tile_range = [&](int t, int64_t &i0, int64_t &plo, int64_t &phi) {
        i0 = (I0 + t) * TM;
        plo = i0 - a.kla > j0 - a.kub ? i0 - a.kla : j0 - a.kub;
        phi = i0 + TM - 1 + a.kua < j0 + TN - 1 + a.klb ? i0 + TM - 1 + a.kua : j0 + TN - 1 + a.klb;
        if (plo < 0) plo = 0;
        if (phi > a.k - 1) phi = a.k - 1;
        plo &= ~(int64_t)3;
    };
    auto iter_seek = [&](TileIter &it) {  // move to the first non-empty tile at or after it.t
        it.valid = false;
        while (it.t < it.ntiles) {
            int64_t plo;
            tile_range(it.t, it.i0, plo, it.phi);
            if (it.phi >= plo) {
                it.pc = plo;
                it.valid = true;
                return;
            }
            ++it.t;
        }
    };
    auto iter_next = [&](TileIter &it) {
        it.pc += KC;
        if (it.pc > it.phi) {
            ++it.t;
            iter_seek(it);
        }
    };

    // staging maps (see gbmm_dmma_kernel): pairs along the contiguous direction
    const int a_ii = 2 * (tid & 31), a_p0 = tid >> 5;
    const int b_pp = 2 * (tid & 7), b_j0 = tid >> 3;
    const int64_t a_step = a.lda - 1, b_step = a.ldb - 1;
    const bool a_al = ((reinterpret_cast<uintptr_t>(a.A) & 15u) == 0) && ((a_step & 1) == 0) && ((a.kua & 1) == 0);
    const bool b_al = ((reinterpret_cast<uintptr_t>(a.B) & 15u) == 0) && ((b_step & 1) == 0) && ((a.kub & 1) == 0);
    const bool cols_in = j0 + TN <= a.n;

    auto cp16 = [&](double *dst, const double *src) {
        const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    };
    auto issue = [&](const TileIter &it, int slot) {
        double *sA = smem + slot * STAGE_DBL;
        double *sB = sA + KC * PA;
        const int64_t i0 = it.i0, pc = it.pc;
        const bool p_in = pc + KC <= a.k;
        const int di = (int)(i0 - pc), dj = (int)(pc - j0);
        // A: rows i0 + a_ii (+1), p = pc + a_p0 + 4q;  A[i,p] = Ad[(kua + i) + (lda-1)*p]
        const double *ap = a.A + (a.kua + i0 + a_ii) + a_step * (pc + a_p0);
        if (a_al && p_in && i0 + TM <= a.m && di + TM - 1 <= kla && di - (KC - 1) >= -kua) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cp16(sA + (a_p0 + 4 * q) * PA + a_ii, ap + a_step * 4 * q);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int pp = a_p0 + 4 * q;
                const int d0 = di + a_ii - pp;  // i - p for the pair's first row
                const bool pk = pc + pp < a.k;
                const bool ok0 = pk && i0 + a_ii < a.m && d0 <= kla && d0 >= -kua;
                const bool ok1 = pk && i0 + a_ii + 1 < a.m && d0 + 1 <= kla && d0 + 1 >= -kua;
                double *dst = sA + pp * PA + a_ii;
                const double *src = ap + a_step * 4 * q;
                if (ok0 && ok1 && a_al) cp16(dst, src);
                else {
                    cp_async8_zfill(dst, ok0 ? src : a.A, ok0);
                    cp_async8_zfill(dst + 1, ok1 ? src + 1 : a.A, ok1);
                }
            }
        }
        // B: p = pc + b_pp (+1), j = j0 + b_j0 + 16q;  B[p,j] = Bd[(kub + p) + (ldb-1)*j]
        const double *bp = a.B + (a.kub + pc + b_pp) + b_step * (j0 + b_j0);
        if (b_al && p_in && cols_in && dj + KC - 1 <= klb && dj - (TN - 1) >= -kub) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cp16(sB + (b_j0 + 16 * q) * PB + b_pp, bp + b_step * 16 * q);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int jj = b_j0 + 16 * q;
                const int d0 = dj + b_pp - jj;  // p - j for the pair's first p
                const bool jk = j0 + jj < a.n;
                const bool ok0 = jk && pc + b_pp < a.k && d0 <= klb && d0 >= -kub;
                const bool ok1 = jk && pc + b_pp + 1 < a.k && d0 + 1 <= klb && d0 + 1 >= -kub;
                double *dst = sB + jj * PB + b_pp;
                const double *src = bp + b_step * 16 * q;
                if (ok0 && ok1 && b_al) cp16(dst, src);
                else {
                    cp_async8_zfill(dst, ok0 ? src : a.B, ok0);
                    cp_async8_zfill(dst + 1, ok1 ? src + 1 : a.B, ok1);
                }
            }
        }
    };

    TileIter P;  // producer
    P.t = 0;
    P.ntiles = ntiles;
    iter_seek(P);
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (P.valid) {
            issue(P, s);
            iter_next(P);
        }
        cp_async_commit();
    }

    const int fr = lane >> 2, fc = lane & 3;
    const int fa_off = fc * PA + wi + fr;
    const int fb_off = (wj + fr) * PB + fc;
    int g = 0;  // consumed chunks so far (ring position)
    for (int t = 0; t < ntiles; ++t) {
        int64_t i0, plo, phi;
        tile_range(t, i0, plo, phi);
        double acc[4][4][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
        for (int64_t pc = plo; pc <= phi; pc += KC, ++g) {
            cp_async_wait<NSTAGE - 2>();
            __syncthreads();
            if (P.valid) {
                issue(P, (g + NSTAGE - 1) % NSTAGE);
                iter_next(P);
            }
            cp_async_commit();
            const double *sA = smem + (g % NSTAGE) * STAGE_DBL;
            const double *sB = sA + KC * PA;
            const int ia0 = (int)(i0 + wi - pc), jb0 = (int)(j0 + wj - pc);  // relative to the chunk start
            const bool full = (ia0 + 31 <= kla) && (ia0 - (KC - 1) >= -kua) && ((KC - 1) - jb0 <= klb) && (-(jb0 + 31) >= -kub);
#pragma unroll
            for (int ks = 0; ks < KC / 4; ++ks) {
                double fa[4], fb[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    fa[x] = sA[fa_off + 4 * ks * PA + 8 * x];
                    fb[x] = sB[fb_off + 8 * x * PB + 4 * ks];
                }
                if (full) {
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], fa[x], fb[y]);
                } else {
                    const int p = 4 * ks;  // relative to the chunk start
                    unsigned ma = 0, mb = 0;
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int ia = ia0 + 8 * x;
                        if (ia - (p + 3) <= kla && (ia + 7) - p >= -kua) ma |= 1u << x;
                        const int jb = jb0 + 8 * x;
                        if ((p + 3) - jb >= -kub && p - (jb + 7) <= klb) mb |= 1u << x;
                    }
                    if (ma != 0 && mb != 0) {
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < 4; ++y)
                                if ((ma >> x & 1u) && (mb >> y & 1u)) dmma884(acc[x][y][0], acc[x][y][1], fa[x], fb[y]);
                    }
                }
            }
        }
        // epilogue of tile t
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
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------
// warp-specialised variant for interior tiles: a producer warp feeds the cp.async ring, eight consumer warps
// (32x16 blocks of the 64x64 tile) drift through it independently
// ------------------------------------------------------------------------------------------
// Why: in gbmm_dmma_kernel the four warps of a CTA meet at a __syncthreads every chunk although their DMMA counts
// differ widely in the ramp chunks (24 % of warp time is barrier stall, profiles/ncu_C4_r01.md).  Here the ring is
// guarded per stage by a `full` mbarrier (armed by the producer lanes' cp.async arrivals) and an `empty` mbarrier
// (one arrival per consumer warp), so a warp that has nothing to do in a chunk runs ahead by up to NSTAGE-1 chunks
// instead of waiting, and twice as many (smaller) warps keep the DMMA pipe fed.
constexpr int WS_CONSUMERS = 8, WS_NT = 32 * (WS_CONSUMERS + 1);

__device__ __forceinline__ void mbar_init_(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_(uint64_t *bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_cp_async_arrive_(uint64_t *bar) {   // arrives once this thread's prior cp.asyncs landed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t *bar, uint32_t parity, unsigned sleep_ns = 0) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0, spins = 0;
    while (!done) {
        // experiment (gbmm_tile 2xx / 3xx): back off between polls so a waiting warp does not take issue slots from the
        // DMMA-issuing warps of its sub-core (benchmarks/micro/fp64_pipes.cu "contention": a DMMA holds the issue port
        // for ~5 of its 16 cycles, other warps' instructions fit in the remaining ~11)
        if (sleep_ns && spins) __nanosleep(sleep_ns);
        if (++spins > (1u << 22)) __trap();   // a lost arrival must surface as a CUDA error, not a hang
        // suspend-time hint: the warp sleeps in hardware (up to ~4 us) instead of burning issue slots that the
        // DMMA-issuing warps of the same sub-core need (the plain spin loop was 31 % of all executed instructions)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity), "r"(4000u)
            : "memory");
    }
}

template <int NSTAGE, int MINB, int NCONS>
__global__ void __launch_bounds__(32 * (NCONS + 1), MINB) gbmm_dmma_ws_kernel(const DmmaArgs a) {
    // NCONS consumer warps: 8 -> 32x16 blocks (8 accumulators per warp), 4 -> 32x32 blocks (16 accumulators, 8 fragment
    // loads per 16 DMMAs instead of 6 per 8, half the per-chunk classification / barrier instructions per DMMA)
    constexpr int WN = 128 / NCONS, YB = WN / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    __shared__ __align__(8) uint64_t bar_full[NSTAGE], bar_empty[NSTAGE];
    // Per stage and consumer warp: the chunk's class, computed ONCE by the producer (lane w classifies for consumer warp w,
    // all warps in one SIMT pass) instead of ~95 instructions in every consumer's chunk loop.  0 = nothing in band for that
    // warp, 0xffffffff = wholly in band (unmasked path), else mA | mB << 16 (block masks of the ramp chunks).  Round 2
    // ablations (profiles/dmma_r02_notes.md): with NO operand copies and NO C stores the kernel still took 5.97 of
    // 6.48 ms -- the time is the consumers' own per-chunk instruction stream, not memory.
    __shared__ unsigned chunk_desc[NSTAGE][8];
    __shared__ unsigned chunk_kind[NSTAGE][8];   // 0 empty, 1 ramp (masked path), 2 wholly in band; a ramp chunk may have ALL block bits set
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    const unsigned sleep_ns = (a.flags & 8) ? 32u : ((a.flags & 16) ? 128u : 0u);
    // a CTA owns `tpc` consecutive row tiles of one 64-column block; the ring runs straight across tile boundaries
    // (the producer is already fetching the next tile while the consumers finish and store the current one)
    const int tpc = a.tpc;
    const int64_t groups = (a.gy + tpc - 1) / tpc;
    // work items = (column block, group of tpc row tiles); a CTA takes every (gridDim)-th item, so a grid of SMs x MINB
    // CTAs is persistent (the ring keeps running from one column block into the next) and a grid of `total` CTAs is not
    const int64_t lin0 = (int64_t)blockIdx.x + (int64_t)gridDim.x * blockIdx.y;
    const int64_t nct = (int64_t)gridDim.x * gridDim.y;
    const int64_t total = ((a.n + TN - 1) / TN) * groups;
    int64_t j0 = 0, by0 = 0, I0 = 0;
    int rot = 0;   // flag 32: block c starts at row tile (-c mod tpc), so neighbouring column blocks -- which run at the same
                   // time -- are on the SAME rows of A at the same step and share those tiles in L2 instead of re-reading them
    auto set_item = [&](int64_t lin) {
        j0 = (lin / groups) * TN;
        by0 = (lin % groups) * tpc;
        I0 = floordiv(j0 - a.kuc, TM);
        if (I0 < 0) I0 = 0;
        rot = (a.flags & 32) ? (int)((tpc - (lin / groups) % tpc) % tpc) : 0;
    };

    // per-tile parameters; valid = interior tile with work (the corner tiles are left to gbmm_dmma_kernel MODE 2)
    auto tile_params = [&](int t, int64_t &i0, int64_t &plo, int &nch) -> bool {
        const int64_t by = by0 + t;
        if (by >= a.gy) return false;
        i0 = (I0 + by) * TM;
        if (i0 >= a.m || i0 > j0 + TN - 1 + a.klc) return false;
        plo = i0 - a.kla > j0 - a.kub ? i0 - a.kla : j0 - a.kub;
        int64_t phi = i0 + TM - 1 + a.kua < j0 + TN - 1 + a.klb ? i0 + TM - 1 + a.kua : j0 + TN - 1 + a.klb;
        if (plo < 0) plo = 0;
        if (phi > a.k - 1) phi = a.k - 1;
        plo &= ~(int64_t)3;
        nch = phi >= plo ? (int)((phi - plo + KC) / KC) : 0;
        return nch > 0 && i0 + TM <= a.m && j0 >= 4 && j0 + TN + 4 <= a.n && plo >= 4 && plo + (int64_t)nch * KC + 4 <= a.k &&
               a.lda >= 33 && a.ldb >= 33 && ((a.lda - 1) & 1) == 0 && ((a.ldb - 1) & 1) == 0 &&
               ((reinterpret_cast<uintptr_t>(a.A + a.kua + i0 + (a.lda - 1) * plo) & 15u) == 0) &&
               ((reinterpret_cast<uintptr_t>(a.B + a.kub + plo + (a.ldb - 1) * j0) & 15u) == 0);
    };

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init_(&bar_full[s], 33);             // the 32 producer lanes' cp.async arrivals + lane 0's release of the descriptors
            mbar_init_(&bar_empty[s], NCONS);  // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int gi = 0;   // chunks issued / consumed so far by this CTA (ring position; identical in every warp)
    if (warp == NCONS) {
        // ---------------- producer: raw dense windows, unpredicated 16-byte copies ----------------
        // A window 64 rows x 16 p: lane owns the row pair 2*lane, one copy per p.  B window 16 p x 64 j: lane owns
        // the p pair 2*(lane & 7) of columns (lane >> 3) + 4q.
        const int64_t a_step = a.lda - 1, b_step4 = 4 * (a.ldb - 1);
        for (int64_t lin = lin0; lin < total; lin += nct)
        for (int tt = 0; tt < tpc; ++tt) {
            if (tt == 0) set_item(lin);
            const int t = tt + rot < tpc ? tt + rot : tt + rot - tpc;
            int64_t i0, plo;
            int nch;
            if (!tile_params(t, i0, plo, nch)) continue;
            const double *a_src = a.A + (a.kua + i0 + 2 * lane) + (a.lda - 1) * plo;
            const double *b_src = a.B + (a.kub + plo + 2 * (lane & 7)) + (a.ldb - 1) * (j0 + (lane >> 3));
            // lane w < NCONS classifies for consumer warp w: rows 32*(w&1).., columns WN*(w>>1)..
            const int c_ia0 = (int)(i0 + (lane & 1) * 32 - plo), c_jb0 = (int)(j0 + (lane >> 1) * WN - plo);
            const int p_kla = (int)a.kla, p_kua = (int)a.kua, p_klb = (int)a.klb, p_kub = (int)a.kub;
            for (int it = 0; it < nch; ++it, ++gi) {
                const int s = gi % NSTAGE;
                if (gi >= NSTAGE) mbar_wait_(&bar_empty[s], (uint32_t)((gi / NSTAGE - 1) & 1), sleep_ns);
                if (lane < NCONS) {
                    const int pr = it * KC;
                    const bool full = (c_ia0 + 31 - pr <= p_kla) && (c_ia0 - (pr + KC - 1) >= -p_kua) && ((pr + KC - 1) - c_jb0 <= p_klb) &&
                                      (pr - (c_jb0 + WN - 1) >= -p_kub);
                    unsigned d = 0xffffffffu, kind = 2u;
                    if (!full) {
                        const int da = c_ia0 - pr, db = c_jb0 - pr;
                        const unsigned mA = band_mask(-p_kua - 7 - da, p_kla + 3 - da);
                        const unsigned mB = band_mask(-p_klb - 7 - db, p_kub + 3 - db) & (YB == 2 ? 0x3333u : 0xffffu);
                        d = mA | (mB << 16);
                        kind = (mA == 0 || mB == 0) ? 0u : 1u;
                    }
                    chunk_desc[s][lane] = d;
                    chunk_kind[s][lane] = kind;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_(&bar_full[s]);   // release: publishes the descriptors with the stage
                double *sA = smem + s * STAGE_DBL;
                double *sB = sA + KC * PA;
                const uint32_t dA = (uint32_t)__cvta_generic_to_shared(sA + 2 * lane);
                const uint32_t dB = (uint32_t)__cvta_generic_to_shared(sB + (lane >> 3) * PB + 2 * (lane & 7));
                if (!(a.flags & 64)) {   // flag 64 (ablation, wrong results): no copies at all -- what the consumers cost when data never has to arrive
#pragma unroll
                for (int q = 0; q < KC; ++q)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dA + (uint32_t)(q * PA * 8)), "l"(a_src + a_step * q) : "memory");
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dB + (uint32_t)(q * 4 * PB * 8)), "l"(b_src + b_step4 * q) : "memory");
                }
                if (a.flags & 4) {   // experiment: issue every copy twice (2x L2 -> smem traffic, same arithmetic)
#pragma unroll
                    for (int q = 0; q < KC; ++q)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dA + (uint32_t)(q * PA * 8)), "l"(a_src + a_step * q) : "memory");
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dB + (uint32_t)(q * 4 * PB * 8)), "l"(b_src + b_step4 * q) : "memory");
                }
                mbar_cp_async_arrive_(&bar_full[s]);
                a_src += a_step * KC;
                b_src += KC;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        return;
    }

    // ---------------- consumers: warp (wa, wb) owns rows 32*wa.., columns 16*wb.. -------------------
    const int wi = (warp & 1) * 32, wj = (warp >> 1) * WN;
    const int fr = lane >> 2, fc = lane & 3;
    const int kla = (int)a.kla, kua = (int)a.kua, klb = (int)a.klb, kub = (int)a.kub;
    const int fa_off = fc * PA + wi + fr;       // + 4*ks*PA + 8*x
    const int fb_off = (wj + fr) * PB + fc;     // + 8*y*PB + 4*ks
    for (int64_t lin = lin0; lin < total; lin += nct)
    for (int tt = 0; tt < tpc; ++tt) {
        if (tt == 0) set_item(lin);
        const int t = tt + rot < tpc ? tt + rot : tt + rot - tpc;
        int64_t i0, plo;
        int nch;
        if (!tile_params(t, i0, plo, nch)) continue;
        const int ia0 = (int)(i0 + wi - plo), jb0 = (int)(j0 + wj - plo);
        double acc[4][YB][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < YB; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;

        for (int it = 0; it < nch; ++it, ++gi) {
            const int s = gi % NSTAGE;
            const int pr = it * KC;
            // every warp observes every phase of full[s] in order (a parity wait only tells adjacent phases apart, so
            // a warp must not skip one), even when the chunk holds nothing for it
            mbar_wait_(&bar_full[s], (uint32_t)((gi / NSTAGE) & 1), sleep_ns);
            const unsigned desc = chunk_desc[s][warp];      // the producer's classification of this chunk for this warp
            unsigned kind = chunk_kind[s][warp];
            if ((a.flags & 256) && kind == 1u) kind = 0u;   // ablation (wrong results): ramp chunks skipped
            if ((a.flags & 512) && kind == 1u) kind = 2u;   // ablation (wrong results): ramp chunks through the unmasked path
            const bool full = kind == 2u;
            const unsigned mA = desc & 0xffffu, mB = desc >> 16;
            if (kind != 0u) {
                const double *sA = smem + s * STAGE_DBL;
                const double *sB = sA + KC * PA;
                if (full) {
#pragma unroll
                    for (int ks = 0; ks < KC / 4; ++ks) {
                        double ga[4], gb[YB];
#pragma unroll
                        for (int x = 0; x < 4; ++x) ga[x] = sA[fa_off + 4 * ks * PA + 8 * x];
#pragma unroll
                        for (int y = 0; y < YB; ++y) gb[y] = sB[fb_off + 8 * y * PB + 4 * ks];
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < YB; ++y) dmma884(acc[x][y][0], acc[x][y][1], ga[x], gb[y]);
                    }
                } else {
                    // Ramp chunk (a band edge of A and / or B crosses it).  Round-2 ablations (profiles/dmma_r02_notes.md): the
                    // former path -- 16 individually predicated DMMAs per k-step, each behind its own bit test and warp
                    // reconvergence -- cost MORE sub-core time with its 24 useful DMMAs per chunk (1.78 ms of the 6.27) than
                    // running all 64 unmasked (1.61 ms).  Now: out-of-band entries are zeroed IN THE FRAGMENTS (per-lane select,
                    // so the column blocks of B need no skipping at all), whole k-steps with nothing in band are skipped, and
                    // the row blocks of A are skipped four DMMAs at a time (one warp-uniform branch per block).
                    const int da = ia0 - pr, db = jb0 - pr;
                    const int va = da + fr - fc + kua, wa = kla + kua;
                    const int vb = -db + fc - fr + kub, wb = klb + kub;
#pragma unroll
                    for (int ks = 0; ks < KC / 4; ++ks) {
                        const unsigned ma = (mA >> (4 * ks)) & 15u, mb = (mB >> (4 * ks)) & (YB == 2 ? 3u : 15u);
                        if (ma == 0 || mb == 0) continue;
                        double gb[YB];
#pragma unroll
                        for (int y = 0; y < YB; ++y) {
                            const double rb = sB[fb_off + 8 * y * PB + 4 * ks];
                            gb[y] = (unsigned)(vb - 8 * y + 4 * ks) <= (unsigned)wb ? rb : 0.0;
                        }
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            if (!(ma >> x & 1u)) continue;
                            const double ra = sA[fa_off + 4 * ks * PA + 8 * x];
                            const double ga = (unsigned)(va + 8 * x - 4 * ks) <= (unsigned)wa ? ra : 0.0;
#pragma unroll
                            for (int y = 0; y < YB; ++y) dmma884(acc[x][y][0], acc[x][y][1], ga, gb[y]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_(&bar_empty[s]);   // this warp is done with stage s (or never needed it)
        }

        // epilogue: C fragment (row = lane/4, cols = 2*(lane%4) + {0,1}) -> band storage; interior tiles need no bounds tests
        if ((a.flags & 128) && acc[0][0][0] != 12345.678) continue;   // flag 128 (ablation, wrong results): no C stores
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < YB; ++y)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t i = i0 + wi + 8 * x + fr;
                    const int64_t j = j0 + wj + 8 * y + 2 * fc + e;
                    const int64_t f = a.kuc + i - j;
                    if (f >= 0 && f <= a.klc + a.kuc) {
                        double *c = a.C + f + a.ldc * j;
                        const double v = a.alpha * acc[x][y][e];
                        *c = (a.beta == 0.0) ? v : fma(a.beta, *c, v);
                    }
                }
    }
}

// ------------------------------------------------------------------------------------------
// warp-autonomous variant for interior tiles (round 2)
// ------------------------------------------------------------------------------------------
// What round 2's ablations showed (profiles/dmma_r02_notes.md): with NO operand copies and NO C stores the warp-specialised
// kernel still needs 5.8 of its 6.2 ms, it runs at 82 % of the DMMA pipe when every chunk is forced through the unmasked
// path and at 62-70 % otherwise -- the loss is the LOCK-STEP of a CTA's four quadrant warps through a shared ring: near both
// ends of a tile's p-range the quadrants sit on different sides of the band edges (full / ramp / nothing), the ring advances
// at the pace of the busiest one, and 19 % of all warp-chunks are pure hand-shakes.
// Here every WARP is its own pipeline, as in gbmm_chain_warp.cu: it owns one 32x32 block of C (block-diagonal order), stages
// ITS OWN dense 32x8 (A) and 8x32 (B) windows with 16-byte cp.async into a private 3-stage ring (no mbarrier, no producer
// warp, __syncwarp only), walks exactly the p-range of its block (no empty chunks), and knows the range of wholly-in-band
// chunks up front (no per-chunk classification on the unmasked path).  A warp's bubbles -- cp.async latency, ramp chunks,
// the epilogue -- are covered by the other eleven warps of the SM, none of which ever waits for it.  The price is twice the
// L2 -> shared-memory traffic per DMMA (windows are not shared between warps), which the "every copy issued twice"
// experiment of round 1 priced at 3 %.
constexpr int WA_KC = 8;                    // p per chunk: two k-steps, 32 DMMAs
constexpr int WA_PA = 36, WA_PB = 12;       // sA[p][i] / sB[j][p] pitches: 2*PA = 8, 2*PB = 24 (mod 32) -> conflict-free fragment loads
constexpr int WA_STAGE = WA_KC * WA_PA + 32 * WA_PB;   // 672 doubles = 5376 B per warp and stage
constexpr int WA_WARPS = 4;

template <int NS>
__global__ void __launch_bounds__(32 * WA_WARPS, 3) gbmm_dmma_wa_kernel(const DmmaArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *ring = reinterpret_cast<double *>(smem_raw) + (size_t)warp * NS * WA_STAGE;
    const int fr = lane >> 2, fc = lane & 3;
    const int kla = (int)a.kla, kua = (int)a.kua, klb = (int)a.klb, kub = (int)a.kub;
    // block-diagonal range of C: blocks (bi, bj = bi - d) with 32 d - 31 <= klc and 32 d + 31 >= -kuc
    const int64_t dlo = -((a.kuc + 31) / 32), dhi = (a.klc + 31) / 32, nd = dhi - dlo + 1;
    const int64_t nbi = a.m / 32, nbj = a.n / 32;
    const int64_t items = nbi * nd;
    const int64_t wid = ((int64_t)blockIdx.x + (int64_t)gridDim.x * blockIdx.y) * WA_WARPS + warp;
    const int64_t nw = (int64_t)gridDim.x * gridDim.y * WA_WARPS;
    // cp.async lane roles: A window 8 p x 32 rows = 8 columns of 256 B; B window 32 columns of 64 B
    const int a_p = lane >> 4, a_r = 2 * (lane & 15);   // + 2q in p
    const int b_j = lane >> 2, b_p = 2 * (lane & 3);    // + 8q in j
    const int fa_off = fc * WA_PA + fr;                 // + 4*ks*PA + 8*x
    const int fb_off = fr * WA_PB + fc;                 // + 8*y*PB + 4*ks

    for (int64_t item = wid; item < items; item += nw) {
        const int64_t bi = item / nd, d = dlo + (item - bi * nd);
        const int64_t bj = bi - d;
        if (bj < 0 || bj >= nbj) continue;
        const int64_t i0 = bi * 32, j0 = bj * 32;
        const int64_t cb = j0 / TN;                      // 64-column block: the corner launch owns some of them entirely
        if (cb < a.cb_lo || cb >= a.cb_hi0) continue;
        int64_t plo = i0 - a.kla > j0 - a.kub ? i0 - a.kla : j0 - a.kub;
        int64_t phi = i0 + 31 + a.kua < j0 + 31 + a.klb ? i0 + 31 + a.kua : j0 + 31 + a.klb;
        if (plo < 0) plo = 0;
        if (phi > a.k - 1) phi = a.k - 1;
        plo &= ~(int64_t)3;
        const int nch = phi >= plo ? (int)((phi - plo + WA_KC) / WA_KC) : 0;
        if (nch == 0) continue;
        const int ia0 = (int)(i0 - plo), jb0 = (int)(j0 - plo);
        // chunks [it_lo, it_hi] are wholly inside both bands (pr = it*KC): the linear inequalities of the chunk classification
        int lo1 = ia0 + 31 - kla, lo2 = jb0 + 31 - kub;
        int hi1 = ia0 + kua - (WA_KC - 1), hi2 = jb0 + klb - (WA_KC - 1);
        int lo = lo1 > lo2 ? lo1 : lo2;
        if (lo < 0) lo = 0;
        const int hi = hi1 < hi2 ? hi1 : hi2;
        const int it_lo = (lo + WA_KC - 1) / WA_KC;
        const int it_hi = hi >= 0 ? hi / WA_KC : -1;

        const double *a_src = a.A + (a.kua + i0 + a_r) + (a.lda - 1) * (plo + a_p);
        const double *b_src = a.B + (a.kub + plo + b_p) + (a.ldb - 1) * (j0 + b_j);
        const int64_t a_step2 = 2 * (a.lda - 1), b_step8 = 8 * (a.ldb - 1);
        auto issue = [&](int it) {
            double *sA = ring + (it % NS) * WA_STAGE;
            double *sB = sA + WA_KC * WA_PA;
            const uint32_t dA = (uint32_t)__cvta_generic_to_shared(sA + a_p * WA_PA + a_r);
            const uint32_t dB = (uint32_t)__cvta_generic_to_shared(sB + b_j * WA_PB + b_p);
            const double *as = a_src + (a.lda - 1) * (int64_t)(it * WA_KC);
            const double *bs = b_src + it * WA_KC;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dA + (uint32_t)(2 * q * WA_PA * 8)), "l"(as + a_step2 * q) : "memory");
#pragma unroll
            for (int q = 0; q < 4; ++q)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dB + (uint32_t)(8 * q * WA_PB * 8)), "l"(bs + b_step8 * q) : "memory");
        };

        double acc[4][4][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;

#pragma unroll
        for (int q = 0; q < NS - 1; ++q) {
            if (q < nch) issue(q);
            cp_async_commit();
        }
        for (int it = 0; it < nch; ++it) {
            if (it + NS - 1 < nch) issue(it + NS - 1);
            cp_async_commit();
            cp_async_wait<NS - 1>();
            __syncwarp();
            const double *sA = ring + (it % NS) * WA_STAGE;
            const double *sB = sA + WA_KC * WA_PA;
            if (it >= it_lo && it <= it_hi) {
#pragma unroll
                for (int ks = 0; ks < WA_KC / 4; ++ks) {
                    double ga[4], gb[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) ga[x] = sA[fa_off + 4 * ks * WA_PA + 8 * x];
#pragma unroll
                    for (int y = 0; y < 4; ++y) gb[y] = sB[fb_off + 8 * y * WA_PB + 4 * ks];
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], ga[x], gb[y]);
                }
            } else {
                // ramp chunk: out-of-band entries zeroed in the fragments, empty k-steps and empty row blocks of A skipped
                const int pr = it * WA_KC;
                const int da = ia0 - pr, db = jb0 - pr;
                const unsigned mA = band_mask(-kua - 7 - da, kla + 3 - da);
                const unsigned mB = band_mask(-klb - 7 - db, kub + 3 - db);
                const int va = da + fr - fc + kua, wa = kla + kua;
                const int vb = -db + fc - fr + kub, wb = klb + kub;
#pragma unroll
                for (int ks = 0; ks < WA_KC / 4; ++ks) {
                    const unsigned ma = (mA >> (4 * ks)) & 15u, mb = (mB >> (4 * ks)) & 15u;
                    if (ma == 0 || mb == 0) continue;
                    double gb[4];
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const double rb = sB[fb_off + 8 * y * WA_PB + 4 * ks];
                        gb[y] = (unsigned)(vb - 8 * y + 4 * ks) <= (unsigned)wb ? rb : 0.0;
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        if (!(ma >> x & 1u)) continue;
                        const double ra = sA[fa_off + 4 * ks * WA_PA + 8 * x];
                        const double ga = (unsigned)(va + 8 * x - 4 * ks) <= (unsigned)wa ? ra : 0.0;
#pragma unroll
                        for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], ga, gb[y]);
                    }
                }
            }
            __syncwarp();   // every lane is done with this stage before the next iteration refills it
        }
        cp_async_wait<0>();
        // epilogue: C fragment (row = lane/4, cols = 2*(lane%4) + {0,1}) -> band storage
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t i = i0 + 8 * x + fr;
                    const int64_t j = j0 + 8 * y + 2 * fc + e;
                    const int64_t f = a.kuc + i - j;
                    if (f >= 0 && f <= a.klc + a.kuc) {
                        double *c = a.C + f + a.ldc * j;
                        const double v = a.alpha * acc[x][y][e];
                        *c = (a.beta == 0.0) ? v : fma(a.beta, *c, v);
                    }
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
    if (beta == 0.0) {
        gbmm_zero_corners_kernel<<<64, 256, 0, h->stream>>>(m, n, klc, kuc, C, ldc);
        LBM_LAUNCH_CHECK(h, "gbmm_zero_corners_kernel");
    }
    a.gy = gy;
    a.flags = 0;
    a.tpc = 1;
    a.cb_lo = gx;      // corner launch covers every column block unless narrowed below
    a.cb_hi0 = gx;
    const int64_t total = gx * gy;
    const int64_t gxx = total < 1048576 ? total : 1048576;
    dim3 grid((unsigned)gxx, (unsigned)((total + gxx - 1) / gxx));
    // gbmm_variant: 0 = default = warp-specialised kernel (producer warp + 4 consumer warps of 32x32, 4-stage ring, 3 CTAs/SM,
    // one CTA per 64-column block walking all its row tiles) for the interior tiles + the corner instantiation of the plain
    // kernel; 20-23: 8 consumer warps of 32x16 with other ring depths / occupancies; 24-28: 4 consumer warps; 29: 8 / 5-stage.  Plain kernel (lean interior + corner instantiations): 3: 2-stage / 5 CTAs,
    // 1: 2-stage / 4 CTAs, 2: 3-stage / 3 CTAs, 4/5: column-persistent kernels, 6: 3-stage / 4 CTAs,
    // 7: 4-stage / 3 CTAs, 8-10: 6 + rotated quadrants / interleaved p-halves, 11: 6 as one launch, 12: 2-stage / 6 CTAs
    // (measured in profiles/configs_r01_*.jsonl; everything lands between 17 and 21 TFLOP/s)
    const int variant = (int)h->tune.gbmm_variant;
    static std::atomic<int> configured_dev_mask[6] = {};
#define LBM_DMMA_LAUNCH(IDX, NS, MB)                                                                         \
    do {                                                                                                      \
        const size_t smem = (size_t)(NS) * STAGE_DBL * sizeof(double);                                        \
        if (!(configured_dev_mask[IDX] & (1 << h->device))) {                                                 \
            LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_kernel<NS, MB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem));                                                     \
            LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_kernel<NS, MB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem));                                                     \
            LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_kernel<NS, MB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem));                                                     \
            configured_dev_mask[IDX] |= (1 << h->device);                                                     \
        }                                                                                                     \
        if (variant == 11) {                                                                                  \
            gbmm_dmma_kernel<NS, MB, 0><<<grid, NT, smem, h->stream>>>(a);                                    \
        } else {                                                                                              \
            gbmm_dmma_kernel<NS, MB, 1><<<grid, NT, smem, h->stream>>>(a);                                    \
            LBM_LAUNCH_CHECK(h, "gbmm_dmma_kernel<interior>");                                                \
            gbmm_dmma_kernel<NS, MB, 2><<<grid, NT, smem, h->stream>>>(a);                                    \
        }                                                                                                     \
    } while (0)
    if (variant == 30 || variant == 31 || variant == 32) {
        // warp-autonomous interior kernel + the corner instantiation of the plain kernel (same corner scan as below)
        auto hfloordiv = [](int64_t x, int64_t y) { return x >= 0 ? x / y : -((-x + y - 1) / y); };
        auto block_all_interior = [&](int64_t cbk) {
            const int64_t j0 = cbk * TN;
            int64_t I0 = hfloordiv(j0 - kuc, TM);
            if (I0 < 0) I0 = 0;
            for (int64_t by = 0; by < gy; ++by) {
                const int64_t i0 = (I0 + by) * TM;
                if (i0 >= m || i0 > j0 + TN - 1 + klc) continue;
                int64_t plo = i0 - kla > j0 - kub ? i0 - kla : j0 - kub;
                int64_t phi = i0 + TM - 1 + kua < j0 + TN - 1 + klb ? i0 + TM - 1 + kua : j0 + TN - 1 + klb;
                if (plo < 0) plo = 0;
                if (phi > k - 1) phi = k - 1;
                plo &= ~(int64_t)3;
                const int64_t nch = phi >= plo ? (phi - plo + KC) / KC : 0;
                const bool ok = nch > 0 && i0 + TM <= m && j0 >= 4 && j0 + TN + 4 <= n && plo >= 4 && plo + nch * KC + 4 <= k &&
                                lda >= 33 && ldb >= 33 && ((lda - 1) & 1) == 0 && ((ldb - 1) & 1) == 0 &&
                                ((reinterpret_cast<uintptr_t>(A + kua + i0 + (lda - 1) * plo) & 15u) == 0) &&
                                ((reinterpret_cast<uintptr_t>(B + kub + plo + (ldb - 1) * j0) & 15u) == 0);
                if (!ok) return false;
            }
            return true;
        };
        const int64_t scan = gx < 64 ? gx : 64;
        int64_t lo = 0, hi = gx;
        while (lo < scan && !block_all_interior(lo)) ++lo;
        bool have_interior = false;
        if (lo < scan) {
            while (hi > lo + 1 && gx - hi < scan && !block_all_interior(hi - 1)) --hi;
            // one more column block goes to the corner kernel at the far end: a 32x32 block's last 8-wide chunk may reach up
            // to 7 columns of p beyond the end of its 64x64 tile's 16-wide chunks
            if (hi > lo + 1 && block_all_interior(hi - 1) && m % 32 == 0 && n % 32 == 0 && (kua & 1) == 0 && (kub & 1) == 0) {
                a.cb_lo = lo;
                a.cb_hi0 = hi - 1;
                have_interior = true;
            }
        }
        if (have_interior) {
            const int64_t saved_lo = a.cb_lo, saved_hi = a.cb_hi0;
            const int64_t dlo = -((kuc + 31) / 32), dhi = (klc + 31) / 32;
            const int64_t items = (m / 32) * (dhi - dlo + 1);
            int64_t ctas = (items + WA_WARPS - 1) / WA_WARPS;
            if (variant == 31 && ctas > (int64_t)h->sm_count * 3) ctas = (int64_t)h->sm_count * 3;   // persistent
            const int64_t gxw = ctas < 1048576 ? ctas : 1048576;
            dim3 grid_wa((unsigned)gxw, (unsigned)((ctas + gxw - 1) / gxw));
            static std::atomic<int> cfg_wa[2] = {};
            if (variant == 32) {
                const size_t smem = (size_t)WA_WARPS * 4 * WA_STAGE * sizeof(double);
                if (!(cfg_wa[1] & (1 << h->device))) {
                    LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_wa_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    cfg_wa[1] |= (1 << h->device);
                }
                // interior blocks only: the kernel skips the corner column blocks [0, cb_lo) and [cb_hi0, gx)
                gbmm_dmma_wa_kernel<4><<<grid_wa, 32 * WA_WARPS, smem, h->stream>>>(a);
            } else {
                const size_t smem = (size_t)WA_WARPS * 3 * WA_STAGE * sizeof(double);
                if (!(cfg_wa[0] & (1 << h->device))) {
                    LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_wa_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    cfg_wa[0] |= (1 << h->device);
                }
                gbmm_dmma_wa_kernel<3><<<grid_wa, 32 * WA_WARPS, smem, h->stream>>>(a);
            }
            LBM_LAUNCH_CHECK(h, "gbmm_dmma_wa_kernel");
            a.cb_lo = saved_lo;
            a.cb_hi0 = saved_hi;
        }
        // corner column blocks (all of them when no interior range was found): the plain zero-filling kernel
        const int64_t total_c = (a.cb_lo + (gx - a.cb_hi0)) * gy;
        const int64_t gcx = total_c < 1048576 ? total_c : 1048576;
        const dim3 grid_c((unsigned)(gcx > 0 ? gcx : 1), (unsigned)(total_c > 0 ? (total_c + gcx - 1) / gcx : 1));
        const size_t smemc = (size_t)2 * STAGE_DBL * sizeof(double);
        static std::atomic<int> cfg_corner2{0};
        if (!(cfg_corner2 & (1 << h->device))) {
            LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_kernel<2, 5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemc));
            LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_kernel<2, 5, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemc));
            cfg_corner2 |= (1 << h->device);
        }
        if (have_interior) {
            if (total_c > 0) gbmm_dmma_kernel<2, 5, 2><<<grid_c, NT, smemc, h->stream>>>(a);
        } else {
            gbmm_dmma_kernel<2, 5, 0><<<grid, NT, smemc, h->stream>>>(a);   // every tile through the general kernel
        }
    } else if (variant == 0 || (variant >= 20 && variant <= 29)) {
        // warp-specialised interior kernel + the corner instantiation of the plain kernel
        // 20: 3-stage / 4 CTAs, 21: 3-stage / 3 CTAs, 22: 4-stage / 3 CTAs, 23: 2-stage / 4 CTAs
        static std::atomic<int> cfg_ws[10] = {};
        // gbmm_tile doubles as "row tiles per CTA" here; default: 5 for the default variant, 1 for the explicit ones
        // default: one CTA walks ALL row tiles of its 64-column block (tpc = gy)
        a.tpc = h->tune.gbmm_tile > 0 ? (int)(h->tune.gbmm_tile % 100) : (variant == 0 ? (int)(gy < 64 ? gy : 64) : 1);
        if (a.tpc < 1) a.tpc = 1;
        if (h->tune.gbmm_tile / 100 == 1) a.flags |= 4;   // (+100: double-copy traffic experiment, results unchanged)
        if (h->tune.gbmm_tile / 100 == 2) a.flags |= 8;   // (+200: poll the ring barriers with a 32 ns back-off)
        if (h->tune.gbmm_tile / 100 == 3) a.flags |= 16;  // (+300: 128 ns back-off)
        if (h->tune.gbmm_tile / 100 == 5) a.flags |= 32;  // (+500: rotated row-tile order, see set_item)
        if (h->tune.gbmm_tile / 100 == 6) a.flags |= 64;  // (+600: ablation, no operand copies)
        if (h->tune.gbmm_tile / 100 == 7) a.flags |= 128; // (+700: ablation, no C stores)
        if (h->tune.gbmm_tile / 100 == 8) a.flags |= 64 | 128; // (+800: both)
        if (h->tune.gbmm_tile / 100 == 9) a.flags |= 64 | 128 | 256;   // (+900: no memory, ramp chunks skipped)
        if (h->tune.gbmm_tile / 100 == 10) a.flags |= 64 | 128 | 512;  // (+1000: no memory, ramp chunks unmasked)
        int64_t total_ws = gx * ((gy + a.tpc - 1) / a.tpc);
        // gbmm_tile 4xx: persistent grid (SMs x resident CTAs, each walking every grid-th work item)
        const bool persistent = h->tune.gbmm_tile / 100 == 4;
        if (persistent && total_ws > (int64_t)h->sm_count * 3) total_ws = (int64_t)h->sm_count * 3;
        const int64_t gws = total_ws < 1048576 ? total_ws : 1048576;
        dim3 grid_ws((unsigned)gws, (unsigned)((total_ws + gws - 1) / gws));
#define LBM_WS_LAUNCH(IDX, NS, MB, NC)                                                                                     \
    do {                                                                                                                   \
        const size_t smem = (size_t)(NS) * STAGE_DBL * sizeof(double);                                                     \
        if (!(cfg_ws[IDX] & (1 << h->device))) {                                                                           \
            LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_ws_kernel<NS, MB, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            cfg_ws[IDX] |= (1 << h->device);                                                                               \
        }                                                                                                                  \
        gbmm_dmma_ws_kernel<NS, MB, NC><<<grid_ws, 32 * (NC + 1), smem, h->stream>>>(a);                                   \
    } while (0)
        if (variant == 0) LBM_WS_LAUNCH(6, 4, 3, 4);         // default = variant 26 (measured 21.6 vs 19.9-21.2 for the others in one process)
        else if (variant == 20) LBM_WS_LAUNCH(0, 3, 4, 8);
        else if (variant == 21) LBM_WS_LAUNCH(1, 3, 3, 8);
        else if (variant == 22) LBM_WS_LAUNCH(2, 4, 3, 8);
        else if (variant == 24) LBM_WS_LAUNCH(4, 3, 4, 4);   // 4 consumer warps of 32x32
        else if (variant == 25) LBM_WS_LAUNCH(5, 3, 3, 4);
        else if (variant == 26) LBM_WS_LAUNCH(6, 4, 3, 4);
        else if (variant == 27) LBM_WS_LAUNCH(7, 5, 2, 4);
        else if (variant == 28) LBM_WS_LAUNCH(8, 6, 2, 4);
        else if (variant == 29) LBM_WS_LAUNCH(9, 5, 2, 8);
        else LBM_WS_LAUNCH(3, 2, 4, 8);
#undef LBM_WS_LAUNCH
        LBM_LAUNCH_CHECK(h, "gbmm_dmma_ws_kernel");
        {
            // Which column blocks still hold non-interior tiles?  The interior predicate (the device one, restated) fails only
            // on a prefix and on a suffix of the column blocks, so scan inwards from both ends until a block is all-interior.
            auto hfloordiv = [](int64_t x, int64_t y) { return x >= 0 ? x / y : -((-x + y - 1) / y); };
            auto block_all_interior = [&](int64_t cbk) {
                const int64_t j0 = cbk * TN;
                int64_t I0 = hfloordiv(j0 - kuc, TM);
                if (I0 < 0) I0 = 0;
                for (int64_t by = 0; by < gy; ++by) {
                    const int64_t i0 = (I0 + by) * TM;
                    if (i0 >= m || i0 > j0 + TN - 1 + klc) continue;
                    int64_t plo = i0 - kla > j0 - kub ? i0 - kla : j0 - kub;
                    int64_t phi = i0 + TM - 1 + kua < j0 + TN - 1 + klb ? i0 + TM - 1 + kua : j0 + TN - 1 + klb;
                    if (plo < 0) plo = 0;
                    if (phi > k - 1) phi = k - 1;
                    plo &= ~(int64_t)3;
                    const int64_t nch = phi >= plo ? (phi - plo + KC) / KC : 0;
                    const bool ok = nch > 0 && i0 + TM <= m && j0 >= 4 && j0 + TN + 4 <= n && plo >= 4 && plo + nch * KC + 4 <= k &&
                                    lda >= 33 && ldb >= 33 && ((lda - 1) & 1) == 0 && ((ldb - 1) & 1) == 0 &&
                                    ((reinterpret_cast<uintptr_t>(A + kua + i0 + (lda - 1) * plo) & 15u) == 0) &&
                                    ((reinterpret_cast<uintptr_t>(B + kub + plo + (ldb - 1) * j0) & 15u) == 0);
                    if (!ok) return false;
                }
                return true;
            };
            const int64_t scan = gx < 64 ? gx : 64;
            int64_t lo = 0, hi = gx;                       // corner blocks: [0, lo) and [hi, gx)
            while (lo < scan && !block_all_interior(lo)) ++lo;
            if (lo < scan) {
                while (hi > lo + 1 && gx - hi < scan && !block_all_interior(hi - 1)) --hi;
                if (hi > lo && block_all_interior(hi - 1)) { a.cb_lo = lo; a.cb_hi0 = hi; }
            }
            const int64_t total_c = (a.cb_lo + (gx - a.cb_hi0)) * gy;
            const int64_t gcx = total_c < 1048576 ? total_c : 1048576;
            const dim3 grid_c((unsigned)(gcx > 0 ? gcx : 1), (unsigned)(total_c > 0 ? (total_c + gcx - 1) / gcx : 1));
            const size_t smem = (size_t)2 * STAGE_DBL * sizeof(double);
            static std::atomic<int> cfg_corner{0};
            if (!(cfg_corner & (1 << h->device))) {
                LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_kernel<2, 5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                cfg_corner |= (1 << h->device);
            }
            if (total_c > 0) gbmm_dmma_kernel<2, 5, 2><<<grid_c, NT, smem, h->stream>>>(a);
        }
    } else if (variant == 4 || variant == 5) {
        // column-block persistent kernels: one CTA per 64-column block
        const int64_t gcol = gx < 1048576 ? gx : 1048576;
        dim3 gridc((unsigned)gcol, (unsigned)((gx + gcol - 1) / gcol));
        static std::atomic<int> cfg_col[2] = {};
        if (variant == 4) {
            const size_t smem = (size_t)2 * STAGE_DBL * sizeof(double);
            if (!(cfg_col[0] & (1 << h->device))) {
                LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_col_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                cfg_col[0] |= (1 << h->device);
            }
            gbmm_dmma_col_kernel<2, 4><<<gridc, NT, smem, h->stream>>>(a);
        } else {
            const size_t smem = (size_t)3 * STAGE_DBL * sizeof(double);
            if (!(cfg_col[1] & (1 << h->device))) {
                LBM_CUDA(h, cudaFuncSetAttribute(gbmm_dmma_col_kernel<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                cfg_col[1] |= (1 << h->device);
            }
            gbmm_dmma_col_kernel<3, 3><<<gridc, NT, smem, h->stream>>>(a);
        }
    } else if (variant == 2) LBM_DMMA_LAUNCH(0, 3, 3);
    else if (variant == 3) LBM_DMMA_LAUNCH(2, 2, 5);
    else if (variant == 6) LBM_DMMA_LAUNCH(3, 3, 4);
    else if (variant == 12) LBM_DMMA_LAUNCH(5, 2, 6);
    else if (variant >= 8 && variant <= 10) { a.flags = variant - 7; LBM_DMMA_LAUNCH(3, 3, 4); }
    else if (variant == 7) LBM_DMMA_LAUNCH(4, 4, 3);
    else if (variant == 1) LBM_DMMA_LAUNCH(1, 2, 4);
    else if (variant == 11) LBM_DMMA_LAUNCH(3, 3, 4);   // 3-stage / 4 CTAs as ONE launch over all tiles
    else LBM_DMMA_LAUNCH(2, 2, 5);   // variant 3 (and any unknown value): 2-stage ring at 5 CTAs (20 warps) per SM: 3-stage ring at 4 CTAs (16 warps) per SM -- 4 x (55.5 + 1) KB just fits
#undef LBM_DMMA_LAUNCH
    LBM_LAUNCH_CHECK(h, "gbmm_dmma_kernel");
    return 0;
}
