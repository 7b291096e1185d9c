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

int main() {
    for (int c : {1, 2, 4}) {
        run<16, 0>("dmma884 x16 independent accumulators", c);
        run<0, 32>("dfma x32 independent", c);
        run<16, 16>("dmma x16 + dfma x16 (128 vs 16 flop-units)", c);
        run<8, 64>("dmma x8 + dfma x64 (equal flops)", c);
    }
    return 0;
}
