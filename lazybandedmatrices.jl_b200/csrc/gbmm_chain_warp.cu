// gbmm_chain_warp.cu -- batched fused lazy chain  D_b = (A_b * B_b) * C_b  (BASELINE config 5:
// 4096 independent brand(4096,4096,4,4) Float32 chains), "warp-autonomous" kernel.
//
// Materialises ApplyMatrix(*, A, B, C) (reference README.md:8-26; the pairwise rule of the
// product band is the one asserted at /root/reference/test/test_misc.jl:32-40) for a batch of
// small-band factors.  The reference route is two gbmm! calls per item with the intermediate
// band written to memory; here A*B never leaves the SM.
//
// Design (HBM-bound: 52 values moved per column for 234 FMAs):
//  * The band storage of a factor is [column][band row] with the band row fastest, so ANY run of
//    columns of an item is one contiguous byte range.  A tile's operands arrive as three 1-D bulk
//    TMA copies (cp.async.bulk, SASS UBLKCP) and the result tile leaves as one bulk TMA store:
//    no staging / layout-transform instructions at all (the previous kernels were issue-bound
//    on exactly that index math, profiles/chain_r01_ncu_summary.md).
//  * The smem tiles keep the RAW layout.  A lane owns 4 adjacent columns, i.e. 4*ld contiguous
//    values = `ld` aligned 16-byte vectors, and lanes are `ld` vectors apart.  For odd `ld`
//    (every symmetric band: 9, 17, 25) that stride is odd in units of 16 bytes, so all
//    LDS.128 / STS.128 are bank-conflict-free and each smem value feeds up to 4 FMAs.
//  * Every WARP is its own pipeline (own smem slab, own mbarriers, __syncwarp only -- there is
//    no CTA-wide barrier after setup): it walks tiles of <= 30 column groups persistently,
//    computes the 32 column groups of A*B its tile needs (2 halo groups recomputed), parks them
//    in smem, and contracts with C into registers.  A and B of the NEXT tile are requested as
//    soon as phase 1 has consumed the current ones, C as soon as its coefficients sit in
//    registers, so the copies fly under phase 2; 8 such warps per SM keep ~100 KB in flight.
//  * Masking is BY INDEX (out-of-matrix band slots may hold anything, NaN included): coefficient
//    loads are select-masked, out-of-range operand columns are zero-filled in smem (item edges
//    only), out-of-matrix result slots are stored as 0 -- same convention as the other kernels.
#include "lbm_common.cuh"

using namespace lbm;

namespace {

__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int r4(int x) { return (x + 3) / 4 * 4; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <typename T, int KLA, int KUA, int KLB, int KUB, int KLC, int KUC, int ABBUF>
struct WarpGeom {
    static constexpr int EPV = 16 / (int)sizeof(T);
    static constexpr int LDA = KLA + KUA + 1, LDB = KLB + KUB + 1, LDC = KLC + KUC + 1;
    static constexpr int L01 = KLA + KLB, U01 = KUA + KUB, LD01 = L01 + U01 + 1;
    static constexpr int L3 = L01 + KLC, U3 = U01 + KUC, LDD = L3 + U3 + 1;
    static constexpr int HQ = r4(cmax(KLC, KUC));   // halo columns of A*B (and B) each side of the D tile
    static constexpr int HB = r4(cmax(KLB, KUB));   // extra halo columns of A each side
    static constexpr int GQ = 32;                    // column groups of A*B per tile (one per lane)
    static constexpr int GD = GQ - 2 * HQ / 4;       // max column groups of D per tile
    static constexpr int NQ = 4 * GQ, NP = NQ + 2 * HB, ND = 4 * GD;
    static constexpr int BYTES_A = NP * LDA * (int)sizeof(T), BYTES_B = NQ * LDB * (int)sizeof(T),
                         BYTES_C = ND * LDC * (int)sizeof(T);
    static constexpr int BYTES_AB = NQ * LD01 * (int)sizeof(T), BYTES_D = ND * LDD * (int)sizeof(T);
    // slab: [A, B] x ABBUF, C, R (= A*B, later D)
    static constexpr int BYTES_ABBUF = BYTES_A + BYTES_B;
    static constexpr int OFF_B = BYTES_A, OFF_C = ABBUF * BYTES_ABBUF, OFF_R = OFF_C + BYTES_C;
    static constexpr int BYTES_WARP = OFF_R + cmax(BYTES_AB, BYTES_D);
    static_assert(GD > 0, "bands too wide for the warp-autonomous chain kernel");
};

// vectors of one 4-column group of a tile with leading dimension LD
template <typename T, int LD> struct GroupVecs { static constexpr int N = 4 * LD * (int)sizeof(T) / 16; };

template <typename T, int LD>
__device__ __forceinline__ void load_group(T (&buf)[4 * LD], const T *smem_tile, int group) {
    using V = typename vec16<T>::type;
    constexpr int NV = GroupVecs<T, LD>::N, EPV = 16 / (int)sizeof(T);
    const V *p = reinterpret_cast<const V *>(smem_tile) + group * NV;
#pragma unroll
    for (int k = 0; k < NV; ++k) *reinterpret_cast<V *>(&buf[EPV * k]) = p[k];
}
template <typename T, int LD>
__device__ __forceinline__ void store_group(T *smem_tile, int group, const T (&buf)[4 * LD]) {
    using V = typename vec16<T>::type;
    constexpr int NV = GroupVecs<T, LD>::N, EPV = 16 / (int)sizeof(T);
    V *p = reinterpret_cast<V *>(smem_tile) + group * NV;
#pragma unroll
    for (int k = 0; k < NV; ++k) p[k] = *reinterpret_cast<const V *>(&buf[EPV * k]);
}

// out[w][U_OUT - U_IN + e + f] += in_col(v)[f] * coef[w][e + KU]  for the 4 owned columns w, where the
// operand column at offset v (relative to the lane's first operand group) pairs with owned column w through
// e = v - V0 - KU - w, V0 = HALO - KU.
template <typename T, int LD_IN, int LD_OUT, int LD_CO, int KL, int KU, int HALO, int ROW_SHIFT>
__device__ __forceinline__ void contract(T (&out)[4 * LD_OUT], const T (&coef)[4 * LD_CO], const T *smem_in, int lane_group) {
    constexpr int V0 = HALO - KU;                 // first needed operand column, relative to 4*lane_group
    constexpr int V1 = HALO + 3 + KL;             // last needed
    constexpr int G0 = V0 / 4, G1 = V1 / 4;
#pragma unroll
    for (int g = G0; g <= G1; ++g) {
        __align__(16) T buf[4 * LD_IN];
        load_group<T, LD_IN>(buf, smem_in, lane_group + g);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int vv = 4 * g + cc - V0;
            if (vv < 0 || vv >= 4 + KL + KU) continue;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int e = vv - KU - w;
                if (e < -KU || e > KL) continue;
                const T c = coef[w * LD_CO + e + KU];
#pragma unroll
                for (int f = 0; f < LD_IN; ++f)
                    out[w * LD_OUT + ROW_SHIFT + e + f] = fma(buf[cc * LD_IN + f], c, out[w * LD_OUT + ROW_SHIFT + e + f]);
            }
        }
    }
}

template <typename T, int KLA, int KUA, int KLB, int KUB, int KLC, int KUC, int NW, int ABBUF>
__global__ void __launch_bounds__(32 * NW, 1)
    gbmm_chain3_warp_kernel(int64_t n, int64_t batch, int tiles_per_item, int groups_base, int groups_rem,
                            const T *__restrict__ A, int64_t sA_, const T *__restrict__ B, int64_t sB_,
                            const T *__restrict__ C, int64_t sC_, T *__restrict__ D, int64_t sD_) {
    using G = WarpGeom<T, KLA, KUA, KLB, KUB, KLC, KUC, ABBUF>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[3 * NW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *slab = smem_raw + (size_t)warp * G::BYTES_WARP;
    T *sC = reinterpret_cast<T *>(slab + G::OFF_C);
    T *sR = reinterpret_cast<T *>(slab + G::OFF_R);   // A*B tile, then (aliased) the D tile
    uint64_t *barAB = &bars[3 * warp], *barC = &bars[3 * warp + 2];   // barAB[0..ABBUF)
    auto bufA = [&](int bsel) { return reinterpret_cast<T *>(slab + bsel * G::BYTES_ABBUF); };
    auto bufB = [&](int bsel) { return reinterpret_cast<T *>(slab + bsel * G::BYTES_ABBUF + G::OFF_B); };
    if (lane == 0) {
        mbar_init(&barAB[0], 1);
        mbar_init(&barAB[1], 1);
        mbar_init(barC, 1);
        mbar_fence_init();
    }
    __syncwarp();

    const int64_t total = batch * (int64_t)tiles_per_item;
    const int64_t stride = (int64_t)gridDim.x * NW;
    int64_t tile = (int64_t)blockIdx.x * NW + warp;

    // tile -> (item, first D column group, D column groups)
    auto decode = [&](int64_t k, int64_t &b, int &g0, int &ng) {
        b = k / tiles_per_item;
        const int ti = (int)(k - b * tiles_per_item);
        g0 = ti * groups_base + (ti < groups_rem ? ti : groups_rem);
        ng = groups_base + (ti < groups_rem ? 1 : 0);
    };
    // request A and B of a tile: zero-fill the out-of-range operand columns (item edges), then bulk copies
    auto request_ab = [&](int64_t b, int g0, int ng, int bsel) {
        T *sA = bufA(bsel), *sB = bufB(bsel);
        uint64_t *bar = &barAB[bsel];
        const int nn = (int)n;
        const int q0 = 4 * g0 - G::HQ, p0 = q0 - G::HB;
        const int nq = 4 * ng + 2 * G::HQ, np = nq + 2 * G::HB;        // columns the tile really needs
        const int a_lo = p0 < 0 ? -p0 : 0;                               // local column range inside [0,n)
        const int a_hi = p0 + np > nn ? nn - p0 : np;
        const int b_lo = q0 < 0 ? -q0 : 0;
        const int b_hi = q0 + nq > nn ? nn - q0 : nq;
        if (a_lo > 0 || a_hi < np) {
            for (int e = lane; e < a_lo * G::LDA; e += 32) sA[e] = (T)0;
            for (int e = a_hi * G::LDA + lane; e < np * G::LDA; e += 32) sA[e] = (T)0;
            fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) {
            const uint32_t ba = (uint32_t)((a_hi - a_lo) * G::LDA * (int)sizeof(T));
            const uint32_t bb = (uint32_t)((b_hi - b_lo) * G::LDB * (int)sizeof(T));
            mbar_expect_tx(bar, ba + bb);
            bulk_g2s(sA + a_lo * G::LDA, A + b * sA_ + (int64_t)(p0 + a_lo) * G::LDA, ba, bar);
            bulk_g2s(sB + b_lo * G::LDB, B + b * sB_ + (int64_t)(q0 + b_lo) * G::LDB, bb, bar);
        }
    };
    auto request_c = [&](int64_t b, int g0, int ng) {
        __syncwarp();
        if (lane == 0) {
            const uint32_t bc = (uint32_t)(4 * ng * G::LDC * (int)sizeof(T));
            mbar_expect_tx(barC, bc);
            bulk_g2s(sC, C + b * sC_ + 4 * (int64_t)g0 * G::LDC, bc, barC);
        }
    };

    int64_t b;
    int g0, ng;
    if (tile < total) {
        decode(tile, b, g0, ng);
        request_ab(b, g0, ng, 0);
        request_c(b, g0, ng);
        if (ABBUF == 2 && tile + stride < total) {
            decode(tile + stride, b, g0, ng);
            request_ab(b, g0, ng, 1);
        }
    }
    uint32_t parity = 0;
    for (int it = 0; tile < total; tile += stride, ++it) {
        const int bsel = ABBUF == 2 ? (it & 1) : 0;
        const uint32_t parAB = ABBUF == 2 ? (uint32_t)((it >> 1) & 1) : parity;
        const T *sA = bufA(bsel), *sB = bufB(bsel);
        decode(tile, b, g0, ng);
        const int n32 = (int)n;
        const int j0 = 4 * g0, q0 = j0 - G::HQ;
        // masks are only needed where the tile's operand window or result rows leave [0,n): item edges
        constexpr int EDGE_LO = cmax(G::U3, G::HQ + G::HB), EDGE_HI = cmax(G::L3, G::HQ + G::HB);
        const bool edge = j0 < EDGE_LO || j0 + 4 * ng + EDGE_HI > n32;
        const int64_t nxt = tile + stride;         // next tile (its C is requested in this iteration)
        const int64_t nab = tile + ABBUF * stride;   // tile whose A, B are requested in this iteration
        int64_t nb = 0, ab_b = 0;
        int ng0 = 0, nng = 0, ab_g0 = 0, ab_ng = 0;
        if (nxt < total) decode(nxt, nb, ng0, nng);
        if (nab < total) decode(nab, ab_b, ab_g0, ab_ng);

        // ---- phase 1: lane u computes columns q0 + 4u .. q0 + 4u + 3 of A*B --------------------
        mbar_wait(&barAB[bsel], parAB);
        const bool p1_active = lane < ng + 2 * G::HQ / 4;
        __align__(16) T acc[4 * G::LD01];
#pragma unroll
        for (int i = 0; i < 4 * G::LD01; ++i) acc[i] = (T)0;
        if (p1_active) {
            __align__(16) T bco[4 * G::LDB];
            load_group<T, G::LDB>(bco, sB, lane);
            if (edge) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int q = q0 + 4 * lane + w;
#pragma unroll
                    for (int ep = 0; ep < G::LDB; ++ep) {
                        const int p = q + ep - KUB;
                        if (!(q >= 0 && q < n32 && p >= 0 && p < n32)) bco[w * G::LDB + ep] = (T)0;
                    }
                }
            }
            // A[i,p] (band row f) * B[p,q] -> (A*B)[i,q], band row U01 + i - q = (U01 - KUA) + e + f
            contract<T, G::LDA, G::LD01, G::LDB, KLB, KUB, G::HB, G::U01 - KUA>(acc, bco, sA, lane);
            if (edge) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int q = q0 + 4 * lane + w;
                    if (!(q >= 0 && q < n32)) {
#pragma unroll
                        for (int r = 0; r < G::LD01; ++r) acc[w * G::LD01 + r] = (T)0;
                    }
                }
            }
        }
        // the A*B tile aliases the previous D tile: its bulk store must have finished READING smem
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        if (p1_active) store_group<T, G::LD01>(sR, lane, acc);
        __syncwarp();                                   // A, B consumed by every lane; A*B visible to every lane
        if (nab < total) request_ab(ab_b, ab_g0, ab_ng, bsel);

        // ---- phase 2: lane t computes columns j0 + 4t .. j0 + 4t + 3 of D ------------------------
        mbar_wait(barC, parity);
        const bool p2_active = lane < ng;
        __align__(16) T cco[4 * G::LDC];
        if (p2_active) {
            load_group<T, G::LDC>(cco, sC, lane);
            if (edge) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int j = j0 + 4 * lane + w;
#pragma unroll
                    for (int ep = 0; ep < G::LDC; ++ep) {
                        const int q = j + ep - KUC;
                        if (!(j < n32 && q >= 0 && q < n32)) cco[w * G::LDC + ep] = (T)0;
                    }
                }
            }
        }
        if (nxt < total) request_c(nb, ng0, nng);      // (__syncwarp inside: C consumed by every lane)
        __align__(16) T dacc[4 * G::LDD];
#pragma unroll
        for (int i = 0; i < 4 * G::LDD; ++i) dacc[i] = (T)0;
        if (p2_active) {
            // (A*B)[i,q] (band row f) * C[q,j] -> D[i,j], band row U3 + i - j = (U3 - U01) + e + f
            contract<T, G::LD01, G::LDD, G::LDC, KLC, KUC, G::HQ, G::U3 - G::U01>(dacc, cco, sR, lane);
            if (edge) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int j = j0 + 4 * lane + w;
#pragma unroll
                    for (int r = 0; r < G::LDD; ++r) {
                        const int i = j + r - G::U3;
                        if (!(i >= 0 && i < n32)) dacc[w * G::LDD + r] = (T)0;
                    }
                }
            }
        }
        __syncwarp();                                   // every lane is done reading A*B
        if (p2_active) store_group<T, G::LDD>(sR, lane, dacc);
        fence_proxy_async();                            // generic-proxy writes -> visible to the bulk store
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(D + b * sD_ + (int64_t)j0 * G::LDD, sR, (uint32_t)(4 * ng * G::LDD * (int)sizeof(T)));
            bulk_commit();
        }
        parity ^= 1u;
    }
    if (lane == 0) bulk_wait0();
}

template <typename T, int KLA, int KUA, int KLB, int KUB, int KLC, int KUC, int ABBUF>
int chain3_warp_launch(lbm_handle_t h, int64_t batch, int64_t n, const T *A, int64_t sA, const T *B, int64_t sB,
                       const T *C, int64_t sC, T *D, int64_t sD) {
    using G = WarpGeom<T, KLA, KUA, KLB, KUB, KLC, KUC, ABBUF>;
    constexpr int NWMAX = (227 * 1024 - 256) / G::BYTES_WARP;
    constexpr int NW = NWMAX >= 8 ? 8 : (NWMAX >= 6 ? 6 : (NWMAX >= 4 ? 4 : (NWMAX >= 2 ? 2 : 1)));
    static_assert(NWMAX >= 1, "tile slab exceeds shared memory");
    constexpr size_t smem = (size_t)NW * G::BYTES_WARP;
    auto kern = gbmm_chain3_warp_kernel<T, KLA, KUA, KLB, KUB, KLC, KUC, NW, ABBUF>;
    static std::atomic<int> configured_dev_mask{0};
    if (!(configured_dev_mask & (1 << h->device))) {
        LBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_dev_mask |= (1 << h->device);
    }
    const int64_t ngroups = n / 4;
    const int64_t tpi = (ngroups + G::GD - 1) / G::GD;
    const int base = (int)(ngroups / tpi), rem = (int)(ngroups % tpi);
    const int64_t total = batch * tpi;
    int64_t ctas = (total + NW - 1) / NW;
    if (ctas > h->sm_count) ctas = h->sm_count;
    kern<<<(unsigned)ctas, 32 * NW, smem, h->stream>>>(n, batch, (int)tpi, base, rem, A, sA, B, sB, C, sC, D, sD);
    LBM_LAUNCH_CHECK(h, "gbmm_chain3_warp_kernel");
    return 0;
}

}  // namespace

// Float32 entry used by chain3_impl (gbmm.cu).  *handled = false when the shape/alignment is not covered.
int lbm_chain3_warp_f32(lbm_handle_t h, int64_t batch, int64_t n, const int64_t *kl, const int64_t *ku, const float *A,
                        int64_t sA, const float *B, int64_t sB, const float *C, int64_t sC, float *D, int64_t sD,
                        bool *handled) {
    *handled = false;
    if (n % 4 != 0 || n < 4 || n > (int64_t)1 << 30 || sA % 4 || sB % 4 || sC % 4 || sD % 4) return 0;
    if (!lbm_aligned16(A) || !lbm_aligned16(B) || !lbm_aligned16(C) || !lbm_aligned16(D)) return 0;
    auto all = [&](int64_t l, int64_t u) {
        return kl[0] == l && kl[1] == l && kl[2] == l && ku[0] == u && ku[1] == u && ku[2] == u;
    };
    *handled = true;
    // default: single-buffered A/B (8 warps/SM for (4,4)^3); chain_variant 5: A/B double-buffered, requested a whole
    // tile ahead (6 warps/SM).  Measured equal within noise (0.61-0.64 ms, profiles/configs_r01_final.jsonl): the
    // kernel sits at the copy-like ceiling of a 53 % read / 47 % write stream, not on load latency.
    const bool single = h->tune.chain_variant != 5;
    if (all(4, 4)) return single ? chain3_warp_launch<float, 4, 4, 4, 4, 4, 4, 1>(h, batch, n, A, sA, B, sB, C, sC, D, sD)
                                 : chain3_warp_launch<float, 4, 4, 4, 4, 4, 4, 2>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
    if (all(2, 2)) return single ? chain3_warp_launch<float, 2, 2, 2, 2, 2, 2, 1>(h, batch, n, A, sA, B, sB, C, sC, D, sD)
                                 : chain3_warp_launch<float, 2, 2, 2, 2, 2, 2, 2>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
    if (all(1, 1)) return single ? chain3_warp_launch<float, 1, 1, 1, 1, 1, 1, 1>(h, batch, n, A, sA, B, sB, C, sC, D, sD)
                                 : chain3_warp_launch<float, 1, 1, 1, 1, 1, 1, 2>(h, batch, n, A, sA, B, sB, C, sC, D, sD);
    *handled = false;
    return 0;
}
