// FP64 pipe peak on this B200 (BASELINE.md section 2 asks for a measured DFMA figure; MEASURED_PEAKS.json has none).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/dfma_peak profiles/dfma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma(double* out, double a, double b, int iters) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    for (int wpsm : {4, 8, 16, 32, 64}) {
        const int threads = 256, blocks = sms * wpsm * 32 / threads;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        dfma<8><<<blocks, threads>>>(out, 1.0000001, 1e-9, 100);
        cudaEventRecord(e0);
        dfma<8><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * iters * (double)blocks * threads;
        printf("warps/SM %2d: %.3f ms  %.2f TFLOP/s FP64 (DFMA)\n", wpsm, ms, flops / ms / 1e9);
    }
    return 0;
}
