// lbm_common.cuh -- shared plumbing of liblbm_b200.so (sm_100a only).
//
// Handle state, LAPACK-style status handling, and the Blackwell async-copy primitives
// (mbarrier + 1-D bulk TMA `cp.async.bulk`, SASS UBLKCP) every streaming kernel uses.
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/lbm_b200.h"

#define LBM_VERSION 10200

// internal status: "this shape cannot take the fused-exchange launch" (nothing was launched)
#define LBM_NOT_FUSED (-100000)

struct lbm_tuning {
    // gbmv narrow (streamed tile) kernel
    int64_t gbmv_tile;        // outputs per tile (0 = auto)
    int64_t gbmv_stages;      // smem ring depth (0 = auto)
    int64_t gbmv_ctas_per_sm; // persistent grid = SMs * this (0 = auto)
    int64_t gbmv_skew;        // experiment knob (elements): see GbmvArgs::skew
    int64_t gbmv_force;       // 0 auto, 1 direct, 2 narrow-stream, 3 wide-stream, 4 wide-T warp kernel (no ring)
    // gbmv wide (column-streaming) kernel
    int64_t gbmvw_chunk;      // columns per chunk (0 = auto)
    int64_t gbmvw_stages;
    int64_t gbmvw_ctas_per_sm;
    // tridiagonal kernel
    int64_t tri_ctas_per_sm;  // 0 = auto
    int64_t tri_force_scalar;
    int64_t tri_force_shuffle; // 1 = use the warp-shuffle kernel for mat-vec too
    int64_t tri_tile, tri_stages;
    // kron
    int64_t kron_tr, kron_tc;
    int64_t kron_stages, kron_ctas_per_sm, kron_force_tile;
    // gbmm
    int64_t gbmm_force;       // 0 auto, 1 generic, 2 dmma
    int64_t gbmm_tile;
    int64_t gbmm_variant;     // DMMA kernel: 0 = 3-stage/3 CTAs per SM, 1 = 2-stage/4 CTAs per SM
    int64_t chain_tile;
    int64_t chain_force_generic;
    int64_t exchange_mode;    // multi-GPU ghost exchange: 0 auto (peer mailboxes when connected), 1 force NCCL send/recv
    int64_t pdl;              // programmatic dependent launch of the streaming kernels: 0 auto (OFF, see lbm_launch_pdl), 1 on
    int64_t exchange_fuse;    // 0 auto (exchange fused into the compute launch when the kernel supports it), 2 never fuse
    int64_t halo_timeout_ms;  // wall-clock bound of a ghost-flag wait on the device (0 = default 30 s)
    int64_t chain_variant;    // 0 auto (= 4 when covered), 1 split-column register kernel, 2/3 transposed-smem kernel, 4 warp-autonomous TMA kernel
};

struct lbm_handle_s {
    int device;
    int sm_count;
    int max_smem_optin;
    cudaStream_t own_stream;
    cudaStream_t stream;
    int64_t launches;
    char err[512];
    lbm_tuning tune;
    void *scratch;
    size_t scratch_bytes;
    cudaStream_t scratch_stream;  // stream of the last call that used the scratch buffer (see lbm_scratch)
    int scratch_used;
    cudaEvent_t ev_scratch;
    // device-side fault word (mapped pinned host memory): a ghost-flag wait that timed out sets it; reported -- as a
    // normal return code, never a trap -- by the next sharded call or lbm_sync (dist.cu: lbm_check_fault)
    int *status_host, *status_dev;
    void *stage[4];           // device staging for *_host entry points
    size_t stage_bytes[4];
    // *_host pipeline: copy-in / copy-out streams and their events (created on first use)
    cudaStream_t h2d_stream, d2h_stream;
    cudaEvent_t ev_copy, ev_kern;
    // multi-GPU: one process per GPU; NCCL communicator owned by the handle (dist.cu)
    void *nccl_comm;
    int rank, world;
    cudaStream_t comm_stream;
    cudaEvent_t ev_ready, ev_halo;
    // peer-memory ghost exchange (dist.cu): a mailbox block in THIS GPU's memory that the neighbours write with
    // NVLink peer stores, and the neighbours' blocks mapped through CUDA IPC
    void *peer_local;            // [flags 256 B][from-left slots 0,1][from-right slots 0,1]
    void *peer_remote[2];        // left / right neighbour's block (IPC mapping), NULL at the ends
    size_t peer_slot_bytes;      // capacity of one mailbox slot
    uint64_t peer_epoch;         // exchanges done so far (identical on all ranks: SPMD)
    int peer_on;                 // 1 once lbm_peer_connect succeeded
    int peer_direct;             // 1: single-process multi-device mode (multi.cu) -- peer_remote are plain peer pointers, no IPC
    int has_comm;                // a row partition exists (NCCL communicator and/or direct peers): rank/world are meaningful
};

int lbm_fail_cuda(lbm_handle_t h, cudaError_t e, const char *what);
int lbm_fail_arg(lbm_handle_t h, int argno, const char *what);
int lbm_scratch(lbm_handle_t h, size_t bytes, void **out);
int lbm_check_fault(lbm_handle_t h);   // 0, or 2001 after a device-side ghost wait timed out (clears the word)
int lbm_stage(lbm_handle_t h, int slot, size_t bytes, void **out);

#define LBM_CUDA(h, call)                                          \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return lbm_fail_cuda(h, e__, #call); \
    } while (0)

#define LBM_LAUNCH_CHECK(h, name)                                  \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) return lbm_fail_cuda(h, e__, name); \
        (h)->launches++;                                           \
    } while (0)

struct lbm_device_guard {
    int prev;
    bool changed;
    explicit lbm_device_guard(int dev) : prev(-1), changed(false) {
        cudaGetDevice(&prev);
        if (prev != dev) { cudaSetDevice(dev); changed = true; }
    }
    ~lbm_device_guard() { if (changed) cudaSetDevice(prev); }
};

static inline bool lbm_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#ifdef __CUDACC__
// Launch with programmatic dependent launch allowed: when the PREVIOUS kernel in the stream is one of ours and has executed
// griddepcontrol.launch_dependents, this grid's CTAs may become resident while the previous grid drains -- launch latency, CTA
// scheduling and the barrier / table set-up overlap its tail.  The kernel executes griddepcontrol.wait before it touches any
// global memory a predecessor may have written (x, y, the mailboxes); a predecessor that never signals (any other kernel)
// serialises as usual.
// MEASURED (benchmarks/kron_small_probe.py, small_kernels.py P): no gain at per-rank sizes (37.1 vs 37.6 us) and a 25-30 % LOSS
// at full size (8192^2 Laplacian 195 -> 253 us): the kernels are one-wave persistent grids that rely on the block scheduler
// spreading CTAs evenly over 148 idle SMs; dependents placed as the predecessor's CTAs retire pile up on the SMs that free
// first.  Hence OFF unless tuning knob pdl = 1.
template <typename... KArgs, typename... Args>
cudaError_t lbm_launch_pdl(lbm_handle_t h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = h->tune.pdl == 1 ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

#ifdef __CUDACC__
namespace lbm {

// see lbm_launch_pdl
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool lbm_aligned16_dev(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// expect_tx WITHOUT arriving: lets the producer arm the byte count, issue the copies, write
// any sub-16-byte tails with plain stores, and only then arrive (release) -- so waiters see
// both the async-proxy bytes and the generic-proxy tail.
__device__ __forceinline__ void mbar_add_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t a = smem_u32(bar);
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (!done) {
        // A lost bulk-copy transaction is a programming error on a par with an illegal address: it must surface as a CUDA
        // error, never as a hung GPU.  The bound is WALL-CLOCK (20 s, read every 64 Ki polls), not a spin count that
        // varies with clocks and contention.
        if ((++spins & 0xFFFFu) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 20000000000ull) __trap();
        }
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
// 1-D bulk TMA global -> shared, completion signalled on `bar` (complete_tx::bytes).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Streaming (read-once) global loads / stores that do not pollute L1.
template <typename T> struct vec16;  // 16-byte vector of T
template <> struct vec16<double> { using type = double2; static constexpr int N = 2; };
template <> struct vec16<float> { using type = float4; static constexpr int N = 4; };

__device__ __forceinline__ double2 ld_stream(const double2 *p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(double2 *p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// Issue the copy of `count` elements src[0..count) into smem `dst` on behalf of the CTA:
// the 16-byte-aligned body goes through ONE bulk TMA copy (issued by the calling thread,
// which must be the thread that armed `bar` with the returned byte count), and the < 16-byte
// tail is written with plain stores by the same thread.  `src` must be 16-byte aligned.
// Returns the number of bytes the bulk copy will signal on the barrier.
template <typename T>
__device__ __forceinline__ uint32_t stage_bytes_bulk(int64_t count) {
    constexpr int EPV = 16 / (int)sizeof(T);
    return (uint32_t)((count / EPV) * EPV * (int64_t)sizeof(T));
}
// Round a copy length up to a whole number of 16-byte vectors when the source array extends that far (`avail` =
// elements readable from src): the whole range then goes through the bulk copy and thread 0 never executes the scalar
// tail, whose dependent global load (~1 us) would sit on the critical path of every tile.  The extra elements land in
// the slack the stage buffer has and are never read.  Used by the multi-RHS kernel, where R+1 serialised tails per tile
// were the bottleneck (3.8 -> 5.5 TB/s at l=u=1, 8 right-hand sides); measured A/B on the single-RHS, fused-sum and
// tridiagonal ring kernels it is 5-8 % SLOWER than the tail (one tail per tile hides behind the ring depth), so they keep it.
template <typename T>
__device__ __forceinline__ int64_t stage_round(int64_t count, int64_t avail) {
    constexpr int EPV = 16 / (int)sizeof(T);
    const int64_t r = (count + EPV - 1) / EPV * EPV;
    return r <= avail ? r : count;
}
template <typename T>
__device__ __forceinline__ void stage_issue(T *dst, const T *src, int64_t count, uint64_t *bar) {
    constexpr int EPV = 16 / (int)sizeof(T);
    const int64_t body = (count / EPV) * EPV;
    if (body > 0) bulk_g2s(dst, src, (uint32_t)(body * (int64_t)sizeof(T)), bar);
    for (int64_t e = body; e < count; ++e) dst[e] = src[e];
}

}  // namespace lbm
#endif  // __CUDACC__
