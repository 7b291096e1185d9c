// Microbenchmark: FP64 throughput ceilings on B200 (sm_100a) -- DMMA.8x8x4 alone, DFMA alone, both mixed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu ; run: ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC, int NFMA>
__global__ void __launch_bounds__(128) pipes(double *out, int iters, double a0, double b0) {
    double c[NACC > 0 ? NACC : 1][2];
    double f[NFMA > 0 ? NFMA : 1];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
    for (int i = 0; i < NFMA; ++i) f[i] = threadIdx.x * 1e-9 + i;
    double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NFMA; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    for (int i = 0; i < NFMA; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the real inner loop of the banded product: per k-step 4 + 4 fragments from shared memory (LDS.64, conflict-free
// pitches as in gbmm_dmma.cu), 16 DMMAs on 16 distinct accumulators -- no masks, no barriers, no global traffic
template <int XB, int YB>
__global__ void __launch_bounds__(128) tile_loop(double *out, int iters) {
    __shared__ double sA[16 * 68], sB[64 * 20];
    for (int i = threadIdx.x; i < 16 * 68; i += blockDim.x) sA[i] = 1.0 + i * 1e-9;
    for (int i = threadIdx.x; i < 64 * 20; i += blockDim.x) sB[i] = 1.0 - i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int fa_off = fc * 68 + (warp & 1) * 32 + fr, fb_off = ((warp >> 1) * 32 + fr) * 20 + fc;
    double acc[XB][YB][2];
    for (int x = 0; x < XB; ++x)
        for (int y = 0; y < YB; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double fa[XB], fb[YB];
#pragma unroll
            for (int x = 0; x < XB; ++x) fa[x] = sA[fa_off + 4 * ks * 68 + 8 * x];
#pragma unroll
            for (int y = 0; y < YB; ++y) fb[y] = sB[fb_off + 8 * y * 20 + 4 * ks];
#pragma unroll
            for (int x = 0; x < XB; ++x)
#pragma unroll
                for (int y = 0; y < YB; ++y) dmma884(acc[x][y][0], acc[x][y][1], fa[x], fb[y]);
        }
    }
    double s = 0;
    for (int x = 0; x < XB; ++x)
        for (int y = 0; y < YB; ++y) s += acc[x][y][0] + acc[x][y][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Does a DMMA in flight block the sub-core's issue port for OTHER warps?  CTA of 8 warps: warps 0-3 (one per sub-core)
// run 16 independent DMMAs per iteration, warps 4-7 (the same four sub-cores) run NALU independent integer adds per
// iteration.  If the two overlap, time = max(DMMA time, ALU time); if every DMMA holds the issue port, time = sum.
template <int NALU>
__global__ void __launch_bounds__(256) issue_contention(double *out, int iters, double a0, double b0) {
    const int warp = threadIdx.x >> 5;
    if (warp < 4) {
        double c[16][2];
        for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
        double a = a0 + threadIdx.x * 1e-12, b = b0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
        double s = 0;
        for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        unsigned r[8];
        for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NALU; ++i) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i & 7]) : "r"(it));
        }
        unsigned s = 0;
        for (int i = 0; i < 8; ++i) s += r[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (double)s;
    }
}

template <int NALU>
void run_contention(int ctas_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * 256);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    issue_contention<NALU><<<sms * ctas_per_sm, 256>>>(out, 100, 1.0, 1e-9);
    cudaEventRecord(e0);
    issue_contention<NALU><<<sms * ctas_per_sm, 256>>>(out, iters, 1.0, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = (double)sms * ctas_per_sm * 4 * iters * 16 * 512.0;
    printf("{\"test\": \"issue contention: 4 DMMA warps + 4 ALU warps, %d int adds per 16 DMMAs\", \"ctas_per_sm\": %d, \"ms\": %.3f, "
           "\"dmma_tflops\": %.2f, \"clk_per_iter_at_1.92GHz\": %.0f}\n", NALU, ctas_per_sm, ms, fl / ms / 1e9, ms * 1e-3 * 1.92e9 / iters / ctas_per_sm);
    cudaFree(out);
}

template <int XB, int YB>
void run_tile(const char *name, int ctas_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * 128);
    const int iters = 5000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tile_loop<XB, YB><<<sms * ctas_per_sm, 128>>>(out, 50);
    cudaEventRecord(e0);
    tile_loop<XB, YB><<<sms * ctas_per_sm, 128>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = (double)sms * ctas_per_sm * 4 * iters * 4.0 * XB * YB * 512.0;
    printf("{\"test\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"dmma_tflops\": %.2f}\n", name, ctas_per_sm, ms, fl / ms / 1e9);
    cudaFree(out);
}

template <int NACC, int NFMA>
void run(const char *name, int ctas_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * 128);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    pipes<NACC, NFMA><<<sms * ctas_per_sm, 128>>>(out, 100, 1.0, 1e-9);
    cudaEventRecord(e0);
    pipes<NACC, NFMA><<<sms * ctas_per_sm, 128>>>(out, iters, 1.0, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)sms * ctas_per_sm * 4;
    const double fl_mma = warps * iters * NACC * 512.0, fl_fma = warps * iters * NFMA * 64.0;
    printf("{\"test\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"total_tflops\": %.2f}\n",
           name, ctas_per_sm, ms, fl_mma / ms / 1e9, fl_fma / ms / 1e9, (fl_mma + fl_fma) / ms / 1e9);
    cudaFree(out);
}

template <int XB, int YB>
void run_tile_long(const char *name, int ctas_per_sm, int launches) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tile_loop<XB, YB><<<sms * ctas_per_sm, 128>>>(out, 50);
    for (int rep = 0; rep < 2; ++rep) {
        const int nl = rep == 0 ? 1 : launches;
        cudaEventRecord(e0);
        for (int l = 0; l < nl; ++l) tile_loop<XB, YB><<<sms * ctas_per_sm, 128>>>(out, 5000);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fl = (double)nl * sms * ctas_per_sm * 4 * 5000 * 4.0 * XB * YB * 512.0;
        printf("{\"test\": \"%s\", \"launches\": %d, \"total_ms\": %.1f, \"dmma_tflops\": %.2f}\n", name, nl, ms, fl / ms / 1e9);
    }
    cudaFree(out);
}

int main(int argc, char **argv) {
    if (argc > 1 && argv[1][0] == 'c') {   // "contention": do other warps' instructions steal DMMA issue time?
        for (int c : {1, 2}) {
            run_contention<0>(c);
            run_contention<64>(c);
            run_contention<128>(c);
            run_contention<256>(c);
            run_contention<512>(c);
        }
        return 0;
    }
    if (argc > 1) {   // "long": sustained load (~2 s) vs a single launch
        run_tile_long<4, 4>("tile loop 32x32, 4 CTAs/SM", 4, 200);
        return 0;
    }
    for (int c : {1, 2, 4}) {
        run<16, 0>("dmma884 x16 independent accumulators", c);
        run<0, 32>("dfma x32 independent", c);
        run<16, 16>("dmma x16 + dfma x16 (128 vs 16 flop-units)", c);
        run<8, 64>("dmma x8 + dfma x64 (equal flops)", c);
    }
    for (int c : {1, 2, 4, 5}) {
        run_tile<4, 4>("tile loop: 8 LDS.64 + 16 DMMA per k-step (32x32 warp tile)", c);
        run_tile<4, 2>("tile loop: 6 LDS.64 + 8 DMMA per k-step (32x16 warp tile)", c);
    }
    return 0;
}
