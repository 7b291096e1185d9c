// gbmm_reg.cu -- narrow banded x banded -> banded with the output band kept in registers.
//
// Materialises ApplyMatrix(*, A, B) (reference README.md:8-26; upstream BandedMatrices gbmm! = n tiny gbmv ccalls,
// SURVEY.md 3.3) for bands up to 17 stored diagonals per factor -- the l,u <= 8 regime of the north-star target,
// which is HBM-bound: bytes = w*n*[(la+ua+1) + (lb+ub+1) + (la+lb+ua+ub+1)].
//
// In diagonal-major storage the product is a convolution of band rows:
//     Cdata[r + s, j] += Adata[r, j + s - kub] * Bdata[s, j],   r in [0, NA), s in [0, NB)     (kuc = kua + kub)
// so only the STORED ROW COUNTS NA = kla+kua+1, NB = klb+kub+1 shape the loop nest, never the individual (kl,ku).
// A thread owns two adjacent output columns and their NA+NB-1 accumulators each, all in registers (compile-time
// indices: the loops below are fully unrolled).  The slab column of A at local index 2t+q feeds column 2t with s = q
// and column 2t+1 with s = q-1, so every value of A read from shared memory is used twice and shared-memory
// bandwidth (the limiter of a one-column-per-thread version: one LDS per FMA) stays below the HBM time.
//
// Shared memory holds the slab of A (columns [j0-kub, ...), ONE contiguous range of the band array), the tile of B and,
// on the way out, the tile of C in exactly the layout they have in HBM (column pitch = row count).  Consequences:
//   * interior tiles with tight leading dimensions arrive as two 1-D bulk TMA copies (cp.async.bulk, SASS UBLKCP)
//     and leave as one bulk TMA store -- zero staging instructions; edge tiles / padded lda use element-wise cp.async
//     with zero-fill, which also masks BY INDEX every out-of-matrix slot (brand() fills them with random numbers);
//   * for odd row counts NAT, NBT (symmetric bands), with lanes one column PAIR apart (2*NAT elements), every
//     2-element vector access (LDS.128 for f64, LDS.64 for f32) is bank-conflict-free, and an element (column 2t+q,
//     row r) is vector-aligned exactly when q + r is even -- a compile-time fact, so each column is read as NAT/2
//     vector loads plus one scalar.  Even counts (2, 4, 6, 10, 12, 14; equal on both factors) are aligned when r is
//     even and pay a 2- or 4-way bank conflict on a traffic that is small next to the HBM time; 8 and 16 rows (8-way)
//     stay on the zero-padded 9 / 17 templates.
#include "lbm_common.cuh"

namespace {

using namespace lbm;

constexpr int GR_NT = 128;          // threads per CTA

template <typename T>
struct GrArgs {
    int64_t m, n, k;
    int kla, kua, klb, kub, klc, kuc;
    int pad;   // the slab starts `pad` columns before j0-kub so that its first byte is 16-byte aligned (bulk TMA)
    int64_t lda, ldb, ldc;
    T alpha, beta;
    const T *A;
    const T *B;
    T *C;
};

template <typename T> struct vec2;
template <> struct vec2<double> { using type = double2; };
template <> struct vec2<float> { using type = float2; };

// resident CTAs per SM the register budget is tuned for: the wide instantiations hold 2*(NAT+NBT-1) accumulators
constexpr int gr_min_blocks(int nat, int nbt) { return nat + nbt >= 26 ? 3 : (nat + nbt >= 14 ? 4 : (nat + nbt >= 8 ? 6 : 10)); }

// one element global -> shared, asynchronously (LDGSTS); ok == false writes a zero without touching global memory
template <typename T>
__device__ __forceinline__ void cp_async_elem(uint32_t dst, const T *src, bool ok) {
    const int nbytes = ok ? (int)sizeof(T) : 0;
    if (sizeof(T) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void gr_bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void gr_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void gr_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void gr_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void gr_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// N consecutive elements starting at p[0]; PAR = parity of the element index of p[0] decides the vector alignment:
// 2-element vector loads where aligned, scalars at the ragged ends
template <typename T, int N, int PAR>
__device__ __forceinline__ void load_run(T *dst, const T *p) {
    using V = typename vec2<T>::type;
    if (PAR) dst[0] = p[0];
#pragma unroll
    for (int r = PAR; r + 1 < N; r += 2) {
        const V v = *reinterpret_cast<const V *>(p + r);
        dst[r] = v.x;
        dst[r + 1] = v.y;
    }
    if ((N - PAR) & 1) dst[N - 1] = p[N - 1];
}

// the same with a RUN-TIME count n <= N: pair positions stay compile-time (register indices static), each pair is a vector
// load when it lies wholly inside the run, a scalar at the ragged end, zeros beyond
template <typename T, int N, int PAR>
__device__ __forceinline__ void load_run_rt(T *dst, const T *p, int n) {
    using V = typename vec2<T>::type;
    if (PAR) dst[0] = 0 < n ? p[0] : (T)0;
#pragma unroll
    for (int r = PAR; r + 1 < N; r += 2) {
        if (r + 1 < n) {
            const V v = *reinterpret_cast<const V *>(p + r);
            dst[r] = v.x;
            dst[r + 1] = v.y;
        } else {
            dst[r] = r < n ? p[r] : (T)0;
            dst[r + 1] = (T)0;
        }
    }
    if ((N - PAR) & 1) dst[N - 1] = N - 1 < n ? p[N - 1] : (T)0;
}

template <typename T, int NAT, int NBT>
struct GrGeom {
    static constexpr int NCT = NAT + NBT - 1;
    // a thread owns one column PAIR per sub-tile of 2*GR_NT columns; narrow instantiations take SUB sub-tiles per
    // bulk copy so that a tile stays ~40 KB (per-tile latency, not bandwidth, bounds smaller tiles)
    static constexpr int B256 = 2 * GR_NT * (NAT + NBT + NCT) * (int)sizeof(T);
    static constexpr int SUB = B256 <= 12 * 1024 ? 4 : (B256 <= 24 * 1024 ? 2 : 1);
    static constexpr int TC = 2 * GR_NT * SUB;                     // output columns per tile
    static constexpr int SA_COLS = (TC + NBT + 3 + 3) & ~3;        // slab columns incl. <= 3 alignment columns (multiple of 4: bulk sizes stay 16-byte multiples)
    static constexpr int SA_ELEMS = SA_COLS * NAT;
    static constexpr int SB_ELEMS = TC * NBT;
    static constexpr int SC_ELEMS = TC * NCT;
    // small instantiations keep the C tile in its OWN region so the next tile's operands can be requested while the
    // current tile is still being written out
    static constexpr bool PIPE = (size_t)(SA_ELEMS + SB_ELEMS + SC_ELEMS) * sizeof(T) <= 76 * 1024;
    static_assert(PIPE || SUB == 1, "aliased C tile needs a single sub-tile");
    static constexpr int SM_ELEMS = PIPE ? (SA_ELEMS + SB_ELEMS + SC_ELEMS)
                                         : ((SA_ELEMS + SB_ELEMS) > SC_ELEMS ? (SA_ELEMS + SB_ELEMS) : SC_ELEMS);
    static constexpr size_t SMEM_BYTES = (size_t)SM_ELEMS * sizeof(T) + 16;   // + the mbarrier
};

// RT (run-time row counts): NA <= NAT, NB <= NBT are NOT the template's.  Shared memory then keeps the operands in THEIR
// OWN HBM layout (column pitch NA, NB, NA+NB-1 instead of NAT, NBT, NCT), so tight operands still arrive and leave as 1-D
// bulk TMA copies; the register loop is the template's, with the rows r >= NA (s >= NB) read as zero by predicated scalar
// loads.  This is what every (NA, NB) without an exact instantiation runs: unequal counts, 8 and 16 rows.  Round 1 padded
// those to the template pitch with element-wise cp.async (45-54 % of the copy ceiling).
template <typename T, int NAT, int NBT, int PADODD, int RT>
__global__ void __launch_bounds__(GR_NT, gr_min_blocks(NAT, NBT)) gbmm_reg_kernel(const GrArgs<T> a) {
    using G = GrGeom<T, NAT, NBT>;
    using V = typename vec2<T>::type;
    constexpr int NCT = G::NCT, SA_COLS = G::SA_COLS, GR_TC = G::TC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *sA = reinterpret_cast<T *>(smem_raw + 16);
    T *sB = sA + G::SA_ELEMS;
    T *sC = G::PIPE ? sB + G::SB_ELEMS : sA;                       // !PIPE: aliases sA/sB after the compute phase

    const int tid = threadIdx.x;
    const int NA = a.kla + a.kua + 1, NB = a.klb + a.kub + 1;
    // shared-memory column pitches.  A C pitch that is a multiple of 8 elements would put every lane's accumulator store on
    // the same banks (16-way): one element of skew, and that tile leaves through the element-wise store instead of the bulk one
    const int PA = RT ? NA : NAT, PB = RT ? NB : NBT, PC = RT ? (NA + NB - 1) + (((NA + NB - 1) & 7) == 0 ? 1 : 0) : NCT;
    const int64_t ntiles = (a.n + GR_TC - 1) / GR_TC;
    const bool tight_in = a.lda == PA && a.ldb == PB && lbm_aligned16_dev(a.A) && lbm_aligned16_dev(a.B);
    const int nbc = a.klc + a.kuc + 1;
    const int off = a.kuc - a.kua - a.kub;                         // C band row of accumulator 0
    const bool tight_out = nbc == PC && a.ldc == PC && off == 0 && a.beta == (T)0 && lbm_aligned16_dev(a.C);
    uint32_t parity = 0;

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // tile geometry / eligibility for the bulk-TMA path
    auto tile_info = [&](int64_t tile, int64_t &j0, int &cols, bool &interior, bool &tma_in) {
        j0 = tile * GR_TC;
        cols = (int)((a.n - j0) < GR_TC ? (a.n - j0) : GR_TC);
        const int64_t p0 = j0 - a.kub - a.pad;                     // global column of slab column 0
        // interior tile: every staged slot is inside the matrix, no index tests needed
        interior = cols == GR_TC && p0 - a.kua >= 0 && p0 + SA_COLS - 1 + a.kla < a.m && p0 + SA_COLS - 1 < a.k &&
                   j0 + GR_TC - 1 + a.klb < a.k;
        tma_in = interior && tight_in && ((p0 * PA * (int64_t)sizeof(T)) & 15) == 0;
    };
    auto request = [&](int64_t j0) {                               // thread 0 only
        gr_fence_proxy_async();                                    // earlier generic-proxy accesses of this smem are ordered first
        const uint32_t ba = (uint32_t)(SA_COLS * PA * (int)sizeof(T)), bb = (uint32_t)(GR_TC * PB * (int)sizeof(T));
        mbar_expect_tx(bar, ba + bb);
        bulk_g2s(sA, a.A + (j0 - a.kub - a.pad) * a.lda, ba, bar);
        bulk_g2s(sB, a.B + j0 * a.ldb, bb, bar);
    };

    bool requested = false;                                        // PIPE: this tile's operands were requested a tile ago
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t j0;
        int cols;
        bool interior, tma_in;
        tile_info(tile, j0, cols, interior, tma_in);
        const int64_t p0 = j0 - a.kub - a.pad;
        const T *Ag = a.A + p0 * a.lda;                            // may point before A for p0 < 0: never dereferenced there
        const T *Bg = a.B + j0 * a.ldb;

        if (!G::PIPE) {
            // the previous tile's bulk store must have finished READING sC (= sA/sB) before anything overwrites it
            if (tid == 0) gr_bulk_wait_read0();
            __syncthreads();
        }

        // ---- stage the A slab and the B tile, HBM layout kept ----
        if (tma_in) {
            if (!requested && tid == 0) request(j0);
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {
            // element-wise cp.async (LDGSTS): nothing waits on a load inside the loop; src-size 0 zero-fills masked slots
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            const int lda32 = (int)a.lda, ldb32 = (int)a.ldb;     // <= 2^20 (checked by the host), c < 2^11: 32-bit offsets
            // the NA (NB) stored rows of every column, walked with an incrementally updated (column, row) pair ...
            {
                const int dc = GR_NT / NA, dr = GR_NT - dc * NA;
                int c = tid / NA, r = tid - c * NA;
                for (int e = tid; e < SA_COLS * NA; e += GR_NT) {
                    bool ok = true;
                    if (!interior) {
                        const int64_t p = p0 + c, i = p + r - a.kua;
                        ok = p >= 0 && p < a.k && i >= 0 && i < a.m;
                    }
                    cp_async_elem<T>(sA_u + (uint32_t)((c * PA + r) * (int)sizeof(T)), ok ? Ag + (r + lda32 * c) : a.A, ok);
                    r += dr;
                    c += dc;
                    if (r >= NA) { r -= NA; ++c; }
                }
            }
            {
                const int dc = GR_NT / NB, dr = GR_NT - dc * NB;
                int c = tid / NB, r = tid - c * NB;
                for (int e = tid; e < GR_TC * NB; e += GR_NT) {
                    bool ok = c < cols;
                    if (ok && !interior) {
                        const int64_t p = j0 + c + r - a.kub;
                        ok = p >= 0 && p < a.k;
                    }
                    cp_async_elem<T>(sB_u + (uint32_t)((c * PB + r) * (int)sizeof(T)), ok ? Bg + (r + ldb32 * c) : a.B, ok);
                    r += dr;
                    c += dc;
                    if (r >= NB) { r -= NB; ++c; }
                }
            }
            // ... and zeros in the NAT-NA (NBT-NB) padding rows of the template
            if (!RT && NA < NAT) {
                const int np = NAT - NA, dc = GR_NT / np, dr = GR_NT - dc * np;
                int c = tid / np, r = tid - c * np;
                for (int e = tid; e < SA_COLS * np; e += GR_NT) {
                    sA[c * NAT + NA + r] = (T)0;
                    r += dr;
                    c += dc;
                    if (r >= np) { r -= np; ++c; }
                }
            }
            if (!RT && NB < NBT) {
                const int np = NBT - NB, dc = GR_NT / np, dr = GR_NT - dc * np;
                int c = tid / np, r = tid - c * np;
                for (int e = tid; e < GR_TC * np; e += GR_NT) {
                    sB[c * NBT + NB + r] = (T)0;
                    r += dr;
                    c += dc;
                    if (r >= np) { r -= np; ++c; }
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
        }
        requested = false;

        const bool tma_out = tight_out && cols == GR_TC && j0 - a.kuc >= 0 && j0 + GR_TC - 1 + a.klc < a.m;
        if (G::PIPE) {
            // the previous tile's bulk store must have finished READING sC before it is rewritten
            if (tid == 0) gr_bulk_wait_read0();
            __syncthreads();
        }
#pragma unroll 1
        for (int sub = 0; sub < G::SUB; ++sub) {
            // ---- compute: two columns, NCT accumulators each, everything in registers ----
            T acc0[NCT], acc1[NCT];
#pragma unroll
            for (int t = 0; t < NCT; ++t) acc0[t] = acc1[t] = (T)0;
            if (RT) {
                // run-time pitches: scalar loads, rows beyond the real counts read as zero (compile-time register indices)
                const int cp = 2 * (tid + sub * GR_NT);
                const T *pa = sA + (cp + a.pad) * PA;
                const T *pb = sB + cp * PB;
#pragma unroll
                for (int q = 0; q <= NBT; ++q) {
                    if (q > NB) break;                             // warp-uniform
                    T av[NAT];
                    // element index (cp+pad+q)*NA + r is even <=> odd NA: pad + q + r even; even NA: r even
                    if (RT == 2) {
                        if ((NA & 1) && ((q + PADODD) & 1)) load_run_rt<T, NAT, 1>(av, pa + q * PA, NA);
                        else load_run_rt<T, NAT, 0>(av, pa + q * PA, NA);
                    } else {
#pragma unroll
                        for (int r = 0; r < NAT; ++r) av[r] = r < NA ? pa[q * PA + r] : (T)0;
                    }
                    const T b0 = q < NB ? pb[q] : (T)0;
                    const T b1 = q >= 1 ? pb[PB + q - 1] : (T)0;
                    if (q < NBT) {
#pragma unroll
                        for (int r = 0; r < NAT; ++r) acc0[r + q] = fma(av[r], b0, acc0[r + q]);
                    }
                    if (q >= 1) {
#pragma unroll
                        for (int r = 0; r < NAT; ++r) acc1[r + q - 1] = fma(av[r], b1, acc1[r + q - 1]);
                    }
                }
            } else {
                const int cp = 2 * (tid + sub * GR_NT);            // first column of this thread's pair
                const T *pa = sA + (cp + a.pad) * NAT;             // slab column of (cp, s = 0); column cp + q starts q*NAT further on
                const T *pb = sB + cp * NBT;                       // [B column cp (NBT values) | B column cp+1]
                T b0n = (T)0, b1n = (T)0;                          // second halves of the B vectors, used at odd q
#pragma unroll
                for (int q = 0; q <= NBT; ++q) {
                    T av[NAT];
                    // index (cp+pad+q)*NAT + r is even <=> (pad + q)*NAT + r even: odd NAT -> pad + q + r even, even NAT -> r even
                    if ((NAT & 1) == 0 || ((q + PADODD) & 1) == 0) load_run<T, NAT, 0>(av, pa + q * NAT);
                    else load_run<T, NAT, 1>(av, pa + q * NAT);
                    T b0 = (T)0, b1 = (T)0;
                    // B values of the pair sit at pb[0 .. 2*NBT): column 0 needs pb[q], column 1 needs pb[NBT + q - 1]; they
                    // are fetched as aligned 2-element vectors one step ahead (which step depends on the parity of NBT)
                    if ((q & 1) == 0) {
                        if (q + 1 < NBT) {
                            const V v = *reinterpret_cast<const V *>(pb + q);
                            b0 = v.x;
                            b0n = v.y;
                        } else if (q < NBT) b0 = pb[q];
                    } else b0 = b0n;
                    if (NBT & 1) {
                        if ((q & 1) == 0) {
                            if (q >= 1) {
                                const V w = *reinterpret_cast<const V *>(pb + NBT + q - 1);   // NBT + q - 1 even
                                b1 = w.x;
                                b1n = w.y;
                            } else b1n = pb[NBT];
                        } else b1 = b1n;
                    } else {
                        if (q & 1) {
                            const V w = *reinterpret_cast<const V *>(pb + NBT + q - 1);       // NBT + q - 1 even
                            b1 = w.x;
                            b1n = w.y;
                        } else b1 = b1n;                                                      // (unused at q = 0)
                    }
                    if (q < NBT) {
#pragma unroll
                        for (int r = 0; r < NAT; ++r) acc0[r + q] = fma(av[r], b0, acc0[r + q]);
                    }
                    if (q >= 1) {
#pragma unroll
                        for (int r = 0; r < NAT; ++r) acc1[r + q - 1] = fma(av[r], b1, acc1[r + q - 1]);
                    }
                }
            }
            if (!G::PIPE) __syncthreads();                         // all reads of sA / sB done: sC (aliased) may overwrite them

            // ---- the C tile, HBM layout: [column cp (NCT values) | column cp+1], written as 2-element vectors ----
            if (RT) {
                const T sc = tma_out ? a.alpha : (T)1;
                T *pc = sC + 2 * (tid + sub * GR_NT) * PC;
#pragma unroll
                for (int t = 0; t < NCT; ++t) {
                    if (t < PC) {
                        pc[t] = sc * acc0[t];
                        pc[PC + t] = sc * acc1[t];
                    }
                }
            } else {
                const T sc = tma_out ? a.alpha : (T)1;             // the bulk store cannot scale: do it here
                T *pc = sC + 2 * (tid + sub * GR_NT) * NCT;
#pragma unroll
                for (int i = 0; i < NCT; ++i) {
                    const int k0 = 2 * i, k1 = 2 * i + 1;
                    const T lo = k0 < NCT ? acc0[k0 < NCT ? k0 : 0] : acc1[k0 >= NCT ? k0 - NCT : 0];
                    const T hi = k1 < NCT ? acc0[k1 < NCT ? k1 : 0] : acc1[k1 >= NCT ? k1 - NCT : 0];
                    V v;
                    v.x = sc * lo;
                    v.y = sc * hi;
                    *reinterpret_cast<V *>(pc + 2 * i) = v;
                }
            }
        }
        if (tma_out) gr_fence_proxy_async();                       // generic-proxy writes of sC -> visible to the bulk store
        __syncthreads();                                           // sC complete; every warp is done with sA / sB
        if (G::PIPE) {
            // request the next tile's operands now: they arrive while this tile drains
            const int64_t nxt = tile + gridDim.x;
            if (nxt < ntiles) {
                int64_t nj0;
                int ncols;
                bool nint, ntma;
                tile_info(nxt, nj0, ncols, nint, ntma);
                if (ntma) {
                    requested = true;
                    if (tid == 0) request(nj0);
                }
            }
        }
        T *Cg = a.C + j0 * a.ldc;
        if (tma_out) {
            if (tid == 0) {
                gr_bulk_s2g(Cg, sC, (uint32_t)(GR_TC * PC * (int)sizeof(T)));
                gr_bulk_commit();
            }
        } else {
            const int NC = NA + NB - 1;
            const int total = cols * nbc;
            const int ldc32 = (int)a.ldc;
            const bool rows_inside = j0 - a.kuc >= 0 && j0 + GR_TC - 1 + a.klc < a.m;   // no out-of-matrix slot in this tile
            // flat element index e = c*nbc + row, advanced incrementally (no division in the loop)
            const int dc = GR_NT / nbc, dr = GR_NT - dc * nbc;
            int c = tid / nbc, row = tid - c * nbc;
            for (int e = tid; e < total; e += GR_NT) {
                const int t = row - off;
                bool in = true;
                if (!rows_inside) {
                    const int64_t i = j0 + c + row - a.kuc;
                    in = i >= 0 && i < a.m;
                }
                T v = (T)0;
                if ((unsigned)t < (unsigned)NC && in) v = sC[c * PC + t];
                T *dst = Cg + (row + ldc32 * c);
                if (a.beta == (T)0) *dst = a.alpha * v;
                else if (in) *dst = fma(a.beta, *dst, a.alpha * v);
                row += dr;
                c += dc;
                if (row >= nbc) { row -= nbc; ++c; }
            }
        }
    }
    if (tid == 0) gr_bulk_wait0();                                 // smem must outlive the last bulk store
}

template <typename T, int NAT, int NBT, int RT = 0>
int gr_launch(lbm_handle_t h, const GrArgs<T> &a) {
    using G = GrGeom<T, NAT, NBT>;
    static_assert((G::SA_ELEMS * sizeof(T)) % 16 == 0 && (G::SB_ELEMS * sizeof(T)) % 16 == 0 && (G::SC_ELEMS * sizeof(T)) % 16 == 0,
                  "bulk copy sizes must be multiples of 16 bytes");
    constexpr size_t smem = G::SMEM_BYTES;
    auto kern = (RT != 1 && (a.pad & 1)) ? gbmm_reg_kernel<T, NAT, NBT, RT == 1 ? 0 : 1, RT> : gbmm_reg_kernel<T, NAT, NBT, 0, RT>;
    static std::atomic<int> configured[2] = {};
    if (!(configured[a.pad & 1] & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[a.pad & 1] |= (1 << h->device);
    }
    int per_sm = 1;
    LBM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GR_NT, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (a.n + G::TC - 1) / G::TC;
    int64_t grid = (int64_t)h->sm_count * per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, GR_NT, smem, h->stream>>>(a);
    LBM_LAUNCH_CHECK(h, "gbmm_reg_kernel");
    return 0;
}

// run-time-count variant: vector loads (RT = 2) for f64 except the widest NA < NB pairs, scalar loads (RT = 1) for f32 and
// those (benchmarks/sweep_gbmm.py, profiles/sweep_gbmm_r02.jsonl: scalar wins every f32 shape by 1-18 points; f64 vector
// wins by up to 8 except (11..15, 13..17) with NA < NB, where scalar wins by 1-8)
template <typename T, int NAT, int NBT>
int gr_launch_rt(lbm_handle_t h, const GrArgs<T> &a, int NA, int NB) {
    if (sizeof(T) == 8 && !(NAT == 17 && NBT == 17 && NA >= 11 && NB > NA)) return gr_launch<T, NAT, NBT, 2>(h, a);
    return gr_launch<T, NAT, NBT, 1>(h, a);
}

template <typename T, int NAT>
int gr_dispatch_b(lbm_handle_t h, const GrArgs<T> &a, int NA, int NB) {
    // exact template counts: vector loads, compile-time pitches; anything else: the run-time-count variant (tuning knob
    // gbmm_force = 4 keeps round 1's zero-padded staging for comparison) ...
    // ... except 8 / 16 rows in either factor: with a 64- / 128-byte column pitch every lane of a warp reads the same banks
    // whatever the thread-to-column map (measured 0.22 / 0.14 of the ceiling run-time vs 0.54 / 0.58 padded to 9 / 17)
    const bool pad = h->tune.gbmm_force == 4 || NA % 8 == 0 || NB % 8 == 0;
    if (NB <= 3) return (NA == NAT && NB == 3) || pad ? gr_launch<T, NAT, 3>(h, a) : gr_launch_rt<T, NAT, 3>(h, a, NA, NB);
    if (NB <= 5) return (NA == NAT && NB == 5) || pad ? gr_launch<T, NAT, 5>(h, a) : gr_launch_rt<T, NAT, 5>(h, a, NA, NB);
    if (NB <= 9) return (NA == NAT && NB == 9) || pad ? gr_launch<T, NAT, 9>(h, a) : gr_launch_rt<T, NAT, 9>(h, a, NA, NB);
    return (NA == NAT && NB == 17) || pad ? gr_launch<T, NAT, 17>(h, a) : gr_launch_rt<T, NAT, 17>(h, a, NA, NB);
}

}  // namespace

// Returns 0 and sets *handled when the register kernel covers the shape (both factors <= 17 stored diagonals).
template <typename T>
int lbm_gbmm_reg(lbm_handle_t h, int64_t m, int64_t n, int64_t k, T alpha, int64_t kla, int64_t kua, const T *A, int64_t lda,
                 int64_t klb, int64_t kub, const T *B, int64_t ldb, T beta, int64_t klc, int64_t kuc, T *C, int64_t ldc,
                 bool *handled) {
    *handled = false;
    const int64_t NA = kla + kua + 1, NB = klb + kub + 1;
    if (NA > 17 || NB > 17 || m <= 0 || n <= 0 || k <= 0) return 0;
    if (kuc < kua + kub) return 0;       // C's band clipped by the matrix size (tiny matrices): generic kernel
    if (klc + kuc + 1 > 4096 || lda > (1 << 20) || ldb > (1 << 20) || ldc > (1 << 20)) return 0;
    *handled = true;
    GrArgs<T> a;
    a.m = m; a.n = n; a.k = k;
    a.kla = (int)kla; a.kua = (int)kua; a.klb = (int)klb; a.kub = (int)kub; a.klc = (int)klc; a.kuc = (int)kuc;
    a.lda = lda; a.ldb = ldb; a.ldc = ldc; a.alpha = alpha; a.beta = beta; a.A = A; a.B = B; a.C = C;
    // tiles start at multiples of 256 columns, so (j0 - kub - pad) * NAT * sizeof(T) is a multiple of 16 for every tile
    // once (kub + pad) is a multiple of 2 (f64) / 4 (f32): NAT is odd
    const int al = 16 / (2 * (int)sizeof(T)) == 1 ? 2 : 4;
    a.pad = (al - (int)(kub % al)) % al;
    // equal odd row counts (symmetric bands l = u = 3, 5, 6, 7 on both factors) have exact instantiations: no padding
    // rows, and tight operands take the bulk-TMA path
    if (NA == NB && NA == 7) return gr_launch<T, 7, 7>(h, a);
    if (NA == NB && NA == 11) return gr_launch<T, 11, 11>(h, a);
    if (NA == NB && NA == 13) return gr_launch<T, 13, 13>(h, a);
    if (NA == NB && NA == 15) return gr_launch<T, 15, 15>(h, a);
    // equal EVEN row counts (A*A', triangular bands, bidiagonal-like factors): exact too; their vector alignment depends on
    // the row parity only.  (8 and 16 rows would put 8 lanes on the same banks: they stay on the padded 9 / 17 templates.)
    if (NA == NB && NA == 2) return gr_launch<T, 2, 2>(h, a);
    if (NA == NB && NA == 4) return gr_launch<T, 4, 4>(h, a);
    if (NA == NB && NA == 6) return gr_launch<T, 6, 6>(h, a);
    if (NA == NB && NA == 10) return gr_launch<T, 10, 10>(h, a);
    if (NA == NB && NA == 12) return gr_launch<T, 12, 12>(h, a);
    if (NA == NB && NA == 14) return gr_launch<T, 14, 14>(h, a);
    if (NA <= 3) return gr_dispatch_b<T, 3>(h, a, (int)NA, (int)NB);
    if (NA <= 5) return gr_dispatch_b<T, 5>(h, a, (int)NA, (int)NB);
    if (NA <= 9) return gr_dispatch_b<T, 9>(h, a, (int)NA, (int)NB);
    return gr_dispatch_b<T, 17>(h, a, (int)NA, (int)NB);
}

template int lbm_gbmm_reg<double>(lbm_handle_t, int64_t, int64_t, int64_t, double, int64_t, int64_t, const double *, int64_t,
                                  int64_t, int64_t, const double *, int64_t, double, int64_t, int64_t, double *, int64_t, bool *);
template int lbm_gbmm_reg<float>(lbm_handle_t, int64_t, int64_t, int64_t, float, int64_t, int64_t, const float *, int64_t,
                                 int64_t, int64_t, const float *, int64_t, float, int64_t, int64_t, float *, int64_t, bool *);
