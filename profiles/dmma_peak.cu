// FP64 tensor-core (DMMA, mma.sync m8n8k4 f64) throughput on this B200, alone and together with DFMA,
// to decide whether the batched B^T D B contraction could gain from tensor cores (north_star: only if compute bound).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/dmma_peak profiles/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int NMMA, int NFMA>
__global__ void mixed(double* out, double a, double b, int iters) {
    double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) dmma(c[i & 3], a + i, b);
#pragma unroll
        for (int i = 0; i < NFMA; ++i) x[i & 7] = fma(x[i & 7], a, b);
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

template <int NMMA, int NFMA>
void run(const char* name, int sms, double* out) {
    const int iters = 4000, threads = 256, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mixed<NMMA, NFMA><<<blocks, threads>>>(out, 1.0000001, 1e-9, 10);
    cudaEventRecord(e0);
    mixed<NMMA, NFMA><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * threads / 32;
    const double mma_flops = 2.0 * 8 * 8 * 4 * NMMA * iters * warps;
    const double fma_flops = 2.0 * 32 * NFMA * iters * warps;
    printf("%-28s %.3f ms  DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s  total %.2f\n", name, ms, mma_flops / ms / 1e9,
           fma_flops / ms / 1e9, (mma_flops + fma_flops) / ms / 1e9);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, 8);
    run<8, 0>("DMMA only", sms, out);
    run<0, 16>("DFMA only", sms, out);
    run<8, 16>("DMMA 8 + DFMA 16 per iter", sms, out);
    run<8, 64>("DMMA 8 + DFMA 64 per iter", sms, out);
    return 0;
}
