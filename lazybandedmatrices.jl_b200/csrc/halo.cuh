// halo.cuh -- ghost exchange FUSED into the compute launch (row-sharded mul!, one process or one device per shard).
//
// The only communication on the path is the nearest-neighbour exchange of the max(kl,ku)-wide ghosts of x
// (SURVEY.md 8e).  Round 1 ran it as two extra kernels on a second stream plus two boundary-row launches behind a
// cudaStreamWaitEvent: 4.75 launches per apply, host-launch bound at 8 GPUs (0.02 ms of traffic per apply).  Here the
// exchange lives INSIDE the streaming kernel, so a sharded apply is ONE launch on ONE stream:
//   * CTA 0 first copies this rank's boundary values of x straight into the neighbours' mailboxes (NVLink peer
//     stores), fences at system scope and publishes the epoch number in their flag words (st.release.sys);
//   * every CTA then streams its tiles as usual -- the bulk-TMA copy of a tile's x window reads the (stale) ghost cells
//     of the local buffer like any other element;
//   * the few tiles whose x window overlaps ghost cells are ordered LAST; before computing such a tile the CTA waits for
//     the neighbour's flag (ld.acquire.sys, bounded by wall-clock) and overwrites the ghost part of the window IN SHARED
//     MEMORY from its own mailbox.  By then the neighbour's push -- issued at the start of ITS launch -- has long landed.
// Mailbox slots are double-buffered by epoch parity; because both directions of a neighbour pair signal every epoch
// (the waiter of epoch e+1 cannot pass before the peer has started launch e+1, i.e. finished reading epoch e) no
// acknowledgement round is needed.  The unfused kernels in dist.cu speak the same protocol, so ranks may mix paths.
#pragma once

#include "lbm_common.cuh"

namespace lbm {

struct HaloFuse {
    int on;                          // 0: the kernel ignores everything below
    int elem_bytes;                  // bytes per ghost unit (sizeof(T), or n2*sizeof(T) for grid columns)
    // push, side 0 = to the LEFT neighbour, 1 = to the RIGHT one (flag == NULL: no such neighbour)
    const char *push_src[2];
    char *push_dst[2];
    unsigned long long *push_flag[2];
    int64_t push_bytes[2];
    // pull, side 0 = ghosts on the LEFT of the owned block (written by the left neighbour), 1 = on the right
    const char *pull_src[2];         // local mailbox slot
    const unsigned long long *pull_flag[2];
    int64_t ghost_begin[2], ghost_end[2];   // ghost range in x-index space (units); empty when there is no neighbour
    int nflags;                      // flag words in use per side: 1 for the 1-D kernels, one per row strip for the grid-column exchange
    int wait_only[2];                // a neighbour exists on that side but sends no ghosts (kl or ku == 0): its flag must still
                                     // be awaited once per epoch, or this rank could lap it and overwrite a slot it is reading
    unsigned long long epoch;
    unsigned long long timeout_ns;   // wall-clock bound of the flag wait
    int *status;                     // host-visible fault word (mapped pinned memory): 0 ok, 1 = wait timed out
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Wait until *flag >= epoch, bounded by WALL-CLOCK (a late neighbour -- first-call module load, a rank-0-only CPU leg, a
// profiler replay -- is not a fault; a lost one must not hang the GPU).  On expiry the fault word is set and the wait
// returns false: the kernel finishes with unspecified boundary values and the host reports the fault at the next call.
__device__ __forceinline__ bool halo_wait_flag(const unsigned long long *flag, unsigned long long epoch,
                                               unsigned long long timeout_ns, int *status) {
    if (ld_acquire_sys(flag) >= epoch) return true;
    const unsigned long long t0 = globaltimer_ns();
    unsigned ns = 32;
    for (;;) {
        if (ld_acquire_sys(flag) >= epoch) return true;
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
        if (globaltimer_ns() - t0 > timeout_ns) {
            if (status) *reinterpret_cast<volatile int *>(status) = 1;
            return false;
        }
    }
}

// Copy `nb` bytes to peer memory (all threads of the CTA).
__device__ __forceinline__ void halo_copy(const char *src, char *dst, int64_t nb, int tid, int nthreads) {
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)nb) & 15u) == 0) {
        const int4 *s4 = reinterpret_cast<const int4 *>(src);
        int4 *d4 = reinterpret_cast<int4 *>(dst);
        for (int64_t i = tid; i < nb / 16; i += nthreads) d4[i] = s4[i];
    } else if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)nb) & 3u) == 0) {
        const int *s1 = reinterpret_cast<const int *>(src);
        int *d1 = reinterpret_cast<int *>(dst);
        for (int64_t i = tid; i < nb / 4; i += nthreads) d1[i] = s1[i];
    } else {
        for (int64_t i = tid; i < nb; i += nthreads) dst[i] = src[i];
    }
}

// Called by every thread of the pushing CTA (1-D kernels: the whole payload, flag word 0).  Payload first, then a
// system-scope fence and a CTA barrier, then the flags.
__device__ __forceinline__ void halo_push(const HaloFuse &hf, int tid, int nthreads) {
#pragma unroll
    for (int s = 0; s < 2; ++s)
        if (hf.push_flag[s]) halo_copy(hf.push_src[s], hf.push_dst[s], hf.push_bytes[s], tid, nthreads);
    __threadfence_system();
    __syncthreads();
    // every flag word of the side (nflags > 1 only when this is the unfused push kernel serving a strip-wise waiter);
    // static indexing of the parameter struct (a dynamic index would copy it to local memory)
    if (hf.push_flag[0])
        for (int i = tid; i < hf.nflags; i += nthreads) st_release_sys(hf.push_flag[0] + i, hf.epoch);
    if (hf.push_flag[1])
        for (int i = tid; i < hf.nflags; i += nthreads) st_release_sys(hf.push_flag[1] + i, hf.epoch);
}

// Called by every thread of the pushing CTA after its last tile: keeps the epochs of a neighbour pair in lock-step on the
// sides this rank receives nothing from (see wait_only).
__device__ __forceinline__ void halo_final_wait(const HaloFuse &hf, int tid, int flag_idx = 0) {
    if (tid == 0 && hf.wait_only[0]) halo_wait_flag(hf.pull_flag[0] + flag_idx, hf.epoch, hf.timeout_ns, hf.status);
    if (tid == 1 && hf.wait_only[1]) halo_wait_flag(hf.pull_flag[1] + flag_idx, hf.epoch, hf.timeout_ns, hf.status);
}

// 1-D windows (gbmv / tridiagonal kernels): the stage's x window holds x[w0 .. w1) at sx[0 ..].  If it overlaps ghost
// cells, wait for the owning neighbour and overwrite them in shared memory from the mailbox.  Must be reached by ALL
// threads of the CTA (it synchronises); the overlap test is CTA-uniform.
// Out of line on purpose: it runs for at most three tiles per launch, and inlining it cost the streaming loop 2-4 % (one GPU,
// benchmarks/halo_selftest.py); the callers guard it with a cheap uniform test on the tile number.
template <typename T>
__device__ __noinline__ void halo_patch_1d(const HaloFuse &hf, T *sx, int64_t w0, int64_t w1, int tid, int nthreads) {
    const bool need0 = hf.ghost_end[0] > hf.ghost_begin[0] && w0 < hf.ghost_end[0] && w1 > hf.ghost_begin[0];
    const bool need1 = hf.ghost_end[1] > hf.ghost_begin[1] && w0 < hf.ghost_end[1] && w1 > hf.ghost_begin[1];
    if (!(need0 || need1)) return;
    if (tid == 0 && need0) halo_wait_flag(hf.pull_flag[0], hf.epoch, hf.timeout_ns, hf.status);
    if (tid == 1 && need1) halo_wait_flag(hf.pull_flag[1], hf.epoch, hf.timeout_ns, hf.status);
    __syncthreads();
    if (need0) {
        const int64_t a = w0 > hf.ghost_begin[0] ? w0 : hf.ghost_begin[0];
        const int64_t b = w1 < hf.ghost_end[0] ? w1 : hf.ghost_end[0];
        const T *mb = reinterpret_cast<const T *>(hf.pull_src[0]);
        for (int64_t i = a + tid; i < b; i += nthreads) sx[i - w0] = __ldcg(mb + (i - hf.ghost_begin[0]));
    }
    if (need1) {
        const int64_t a = w0 > hf.ghost_begin[1] ? w0 : hf.ghost_begin[1];
        const int64_t b = w1 < hf.ghost_end[1] ? w1 : hf.ghost_end[1];
        const T *mb = reinterpret_cast<const T *>(hf.pull_src[1]);
        for (int64_t i = a + tid; i < b; i += nthreads) sx[i - w0] = __ldcg(mb + (i - hf.ghost_begin[1]));
    }
    // the patched words are generic-proxy stores into a buffer the async proxy (bulk TMA) overwrites when the stage is
    // re-issued: order the two proxies before the CTA-wide barrier that precedes that re-issue
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
}

}  // namespace lbm

// ---- internal cross-file entry points (not part of the C ABI) --------------------------------------------------------
int lbm_internal_dgbmv_fused(lbm_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, double alpha, const double *A,
                             int64_t lda, const double *x, double beta, double *y, const lbm::HaloFuse *hf);
int lbm_internal_sgbmv_fused(lbm_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, float alpha, const float *A,
                             int64_t lda, const float *x, float beta, float *y, const lbm::HaloFuse *hf);
int lbm_internal_dtridiag_rows_fused(lbm_handle_t h, int64_t n, const double *dl, const double *d, const double *du, double alpha,
                                     const double *x, double beta, double *y, int64_t row_begin, int64_t row_end,
                                     const lbm::HaloFuse *hf);
int lbm_internal_stridiag_rows_fused(lbm_handle_t h, int64_t n, const float *dl, const float *d, const float *du, float alpha,
                                     const float *x, float beta, float *y, int64_t row_begin, int64_t row_end,
                                     const lbm::HaloFuse *hf);
int lbm_internal_exchange(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, void *buf, int64_t unit_bytes, int64_t *plan10,
                          int overlap, lbm::HaloFuse *hf, int *fused, int nflags);
int lbm_internal_exchange_unfuse(lbm_handle_t h, int64_t n, int64_t kl, int64_t ku, void *buf, int64_t unit_bytes, int nflags);
int lbm_internal_comm_resources(lbm_handle_t h);
int lbm_internal_peer_alloc_local(lbm_handle_t h, size_t slot_bytes);
