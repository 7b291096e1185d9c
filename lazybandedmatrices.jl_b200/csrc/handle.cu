// handle.cu -- handle lifetime, status strings, memory helpers, tuning knobs, the
// counter-based synthetic-data generator and the pure-integer structure helper.
#include "lbm_common.cuh"

#include <stdlib.h>

static char g_create_err[512] = "";

int lbm_fail_cuda(lbm_handle_t h, cudaError_t e, const char *what) {
    char *dst = h ? h->err : g_create_err;
    snprintf(dst, 512, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();  // clear sticky-less error state
    return (int)e > 0 ? (int)e : 1;
}

int lbm_fail_arg(lbm_handle_t h, int argno, const char *what) {
    char *dst = h ? h->err : g_create_err;
    snprintf(dst, 512, "argument %d invalid: %s", argno, what);
    return -argno;
}

// One scratch buffer per handle (chain temporaries, Kronecker Z buffers, reduction partials).  The handle's stream can be
// re-bound between calls (the Python layer follows torch's current stream), so two calls issued from different streams
// could overlap on the device and overwrite each other's scratch: when the bound stream differs from the one the
// previous user ran on, the new stream first waits for everything queued on the old one.
int lbm_scratch(lbm_handle_t h, size_t bytes, void **out) {
    lbm_device_guard g(h->device);   // callers driving several devices from one thread (multi.cu) may sit on another device
    if (h->scratch_used && h->scratch_stream != h->stream) {
        if (!h->ev_scratch) LBM_CUDA(h, cudaEventCreateWithFlags(&h->ev_scratch, cudaEventDisableTiming));
        if (cudaEventRecord(h->ev_scratch, h->scratch_stream) == cudaSuccess)
            LBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_scratch, 0));
        else
            cudaGetLastError();   // the old stream no longer exists: its work has drained
    }
    h->scratch_stream = h->stream;
    h->scratch_used = 1;
    if (bytes > h->scratch_bytes) {
        if (h->scratch) LBM_CUDA(h, cudaFree(h->scratch));
        h->scratch = nullptr;
        h->scratch_bytes = 0;
        LBM_CUDA(h, cudaMalloc(&h->scratch, bytes));
        h->scratch_bytes = bytes;
    }
    *out = h->scratch;
    return 0;
}

int lbm_stage(lbm_handle_t h, int slot, size_t bytes, void **out) {
    lbm_device_guard g(h->device);
    if (bytes > h->stage_bytes[slot]) {
        if (h->stage[slot]) LBM_CUDA(h, cudaFree(h->stage[slot]));
        h->stage[slot] = nullptr;
        h->stage_bytes[slot] = 0;
        LBM_CUDA(h, cudaMalloc(&h->stage[slot], bytes));
        h->stage_bytes[slot] = bytes;
    }
    *out = h->stage[slot];
    return 0;
}

int lbm_check_fault(lbm_handle_t h) {
    if (!h || !h->status_host) return 0;
    const int code = *reinterpret_cast<volatile int *>(h->status_host);
    if (!code) return 0;
    *reinterpret_cast<volatile int *>(h->status_host) = 0;
    snprintf(h->err, 512,
             "device fault %d: a ghost-exchange wait timed out (the neighbour rank did not publish its epoch within %lld ms); "
             "the boundary rows of the last sharded result are unspecified",
             code, (long long)(h->tune.halo_timeout_ms > 0 ? h->tune.halo_timeout_ms : 30000));
    return 2000 + code;
}

extern "C" {

int lbm_version(void) { return LBM_VERSION; }

int lbm_create(lbm_handle_t *out, int device) {
    if (!out) return lbm_fail_arg(nullptr, 1, "out == NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return lbm_fail_cuda(nullptr, e, "cudaGetDeviceCount (no CUDA device: this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return lbm_fail_arg(nullptr, 2, "device index out of range");
    lbm_device_guard g(device);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return lbm_fail_cuda(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major < 10) {
        snprintf(g_create_err, 512, "device %d is sm_%d%d; liblbm_b200 ships sm_100a code only", device, prop.major,
                 prop.minor);
        return (int)cudaErrorNoKernelImageForDevice;
    }
    lbm_handle_t h = (lbm_handle_t)calloc(1, sizeof(lbm_handle_s));
    if (!h) return (int)cudaErrorMemoryAllocation;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        free(h);
        return lbm_fail_cuda(nullptr, e, "cudaStreamCreateWithFlags");
    }
    h->stream = h->own_stream;
    h->err[0] = 0;
    h->rank = 0;
    h->world = 1;
    if (cudaHostAlloc((void **)&h->status_host, 64, cudaHostAllocMapped) == cudaSuccess) {
        memset(h->status_host, 0, 64);
        if (cudaHostGetDevicePointer((void **)&h->status_dev, h->status_host, 0) != cudaSuccess) h->status_dev = nullptr;
    } else {
        h->status_host = nullptr;
    }
    cudaGetLastError();
    *out = h;
    return 0;
}

int lbm_destroy(lbm_handle_t h) {
    if (!h) return 0;
    lbm_device_guard g(h->device);
    cudaStreamSynchronize(h->stream);
    lbm_comm_destroy(h);
    if (h->h2d_stream) {
        cudaStreamDestroy(h->h2d_stream);
        cudaStreamDestroy(h->d2h_stream);
        cudaEventDestroy(h->ev_copy);
        cudaEventDestroy(h->ev_kern);
    }
    if (h->scratch) cudaFree(h->scratch);
    if (h->ev_scratch) cudaEventDestroy(h->ev_scratch);
    if (h->status_host) cudaFreeHost(h->status_host);
    for (int i = 0; i < 4; ++i)
        if (h->stage[i]) cudaFree(h->stage[i]);
    cudaStreamDestroy(h->own_stream);
    free(h);
    return 0;
}

int lbm_set_stream(lbm_handle_t h, void *cuda_stream) {
    if (!h) return -1;
    h->stream = (cudaStream_t)cuda_stream;  // NULL is a valid stream: the legacy default stream
    return 0;
}

int lbm_use_own_stream(lbm_handle_t h) {
    if (!h) return -1;
    h->stream = h->own_stream;
    return 0;
}

int lbm_sync(lbm_handle_t h) {
    if (!h) return -1;
    lbm_device_guard g(h->device);
    LBM_CUDA(h, cudaStreamSynchronize(h->stream));
    return lbm_check_fault(h);
}

const char *lbm_last_error_string(lbm_handle_t h) { return h ? h->err : g_create_err; }

int64_t lbm_launch_count(lbm_handle_t h) { return h ? h->launches : -1; }

#define LBM_TUNE_KEYS(X)                                                                                       \
    X(gbmv_tile) X(gbmv_stages) X(gbmv_ctas_per_sm) X(gbmv_force) X(gbmv_skew) X(gbmvw_chunk) X(gbmvw_stages)               \
    X(gbmvw_ctas_per_sm) X(tri_ctas_per_sm) X(tri_force_scalar) X(tri_force_shuffle) X(tri_tile) X(tri_stages) X(kron_tr) X(kron_tc) X(kron_stages) X(kron_ctas_per_sm) X(kron_force_tile) X(gbmm_force) X(gbmm_tile) X(gbmm_variant) \
    X(chain_tile) X(chain_force_generic) X(chain_variant) X(exchange_mode) X(exchange_fuse) X(halo_timeout_ms) X(pdl)

int lbm_set_tuning(lbm_handle_t h, const char *key, int64_t value) {
    if (!h) return -1;
    if (!key) return lbm_fail_arg(h, 2, "key == NULL");
#define X(name)                        \
    if (strcmp(key, #name) == 0) {     \
        h->tune.name = value;          \
        return 0;                      \
    }
    LBM_TUNE_KEYS(X)
#undef X
    return lbm_fail_arg(h, 2, "unknown tuning key");
}

int64_t lbm_get_tuning(lbm_handle_t h, const char *key) {
    if (!h || !key) return -1;
#define X(name) \
    if (strcmp(key, #name) == 0) return h->tune.name;
    LBM_TUNE_KEYS(X)
#undef X
    return -1;
}

// ---- memory helpers ------------------------------------------------------------------
int lbm_malloc(lbm_handle_t h, void **dptr, size_t bytes) {
    if (!h) return -1;
    if (!dptr) return lbm_fail_arg(h, 2, "dptr == NULL");
    lbm_device_guard g(h->device);
    LBM_CUDA(h, cudaMalloc(dptr, bytes ? bytes : 16));
    return 0;
}
int lbm_free(lbm_handle_t h, void *dptr) {
    if (!h) return -1;
    lbm_device_guard g(h->device);
    LBM_CUDA(h, cudaFree(dptr));
    return 0;
}
int lbm_host_alloc(lbm_handle_t h, void **hptr, size_t bytes) {
    if (!h) return -1;
    if (!hptr) return lbm_fail_arg(h, 2, "hptr == NULL");
    lbm_device_guard g(h->device);
    LBM_CUDA(h, cudaMallocHost(hptr, bytes ? bytes : 16));
    return 0;
}
int lbm_host_free(lbm_handle_t h, void *hptr) {
    if (!h) return -1;
    lbm_device_guard g(h->device);
    LBM_CUDA(h, cudaFreeHost(hptr));
    return 0;
}
int lbm_memcpy_h2d(lbm_handle_t h, void *dst, const void *src, size_t bytes) {
    if (!h) return -1;
    lbm_device_guard g(h->device);
    if (bytes) LBM_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
}
int lbm_memcpy_d2h(lbm_handle_t h, void *dst, const void *src, size_t bytes) {
    if (!h) return -1;
    lbm_device_guard g(h->device);
    if (bytes) LBM_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}
int lbm_memcpy_d2d(lbm_handle_t h, void *dst, const void *src, size_t bytes) {
    if (!h) return -1;
    lbm_device_guard g(h->device);
    if (bytes) LBM_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}
int lbm_memset0(lbm_handle_t h, void *dst, size_t bytes) {
    if (!h) return -1;
    lbm_device_guard g(h->device);
    if (bytes) LBM_CUDA(h, cudaMemsetAsync(dst, 0, bytes, h->stream));
    return 0;
}

// ---- structure helper ------------------------------------------------------------------
int lbm_prod_bandwidths(int64_t m, int64_t n, int nf, const int64_t *kl, const int64_t *ku, int64_t *out_kl,
                        int64_t *out_ku) {
    if (m < 0) return -1;
    if (n < 0) return -2;
    if (nf < 1) return -3;
    if (!kl) return -4;
    if (!ku) return -5;
    if (!out_kl) return -6;
    if (!out_ku) return -7;
    int64_t l = 0, u = 0;
    for (int i = 0; i < nf; ++i) {
        l += kl[i];
        u += ku[i];
    }
    *out_kl = l < m - 1 ? l : m - 1;
    *out_ku = u < n - 1 ? u : n - 1;
    return 0;
}

}  // extern "C"

// ---- generator ---------------------------------------------------------------------------
namespace {
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t sm64(uint64_t seed, uint64_t idx) {
    return mix64(seed + (idx + 1ULL) * 0x9E3779B97F4A7C15ULL);
}
__device__ __forceinline__ void gen(uint64_t bits, double &v) { v = (double)(bits >> 11) * 0x1.0p-53; }
__device__ __forceinline__ void gen(uint64_t bits, float &v) { v = (float)(bits >> 40) * 0x1.0p-24f; }

template <typename T>
__global__ void __launch_bounds__(256) fill_uniform_kernel(uint64_t seed, int64_t offset, int64_t count, T *dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        T v;
        gen(sm64(seed, (uint64_t)(offset + i)), v);
        dst[i] = v;
    }
}

template <typename T>
int fill_uniform(lbm_handle_t h, uint64_t seed, int64_t offset, int64_t count, T *dst) {
    if (!h) return -1;
    if (count < 0) return lbm_fail_arg(h, 4, "count < 0");
    if (count == 0) return 0;
    if (!dst) return lbm_fail_arg(h, 5, "dst == NULL");
    lbm_device_guard g(h->device);
    int64_t blocks = (count + 255) / 256;
    int64_t cap = (int64_t)h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    fill_uniform_kernel<T><<<(unsigned)blocks, 256, 0, h->stream>>>(seed, offset, count, dst);
    LBM_LAUNCH_CHECK(h, "fill_uniform_kernel");
    return 0;
}
}  // namespace

extern "C" int lbm_dfill_uniform(lbm_handle_t h, uint64_t seed, int64_t offset, int64_t count, double *dst) {
    return fill_uniform<double>(h, seed, offset, count, dst);
}
extern "C" int lbm_sfill_uniform(lbm_handle_t h, uint64_t seed, int64_t offset, int64_t count, float *dst) {
    return fill_uniform<float>(h, seed, offset, count, dst);
}
