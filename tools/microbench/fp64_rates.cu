// Microbenchmark: FP64 FMA rate of the CUDA cores (DFMA) vs the tensor-core path (mma.sync.m8n8k4.f64, DMMA) on one GPU.
// Decides whether the D-matrix sweeps are worth casting as batched small GEMMs (BASELINE north_star: "only if it gains").
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a[8], x = 1.0000001, y = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
    double c[8][2], a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3;
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 256, blocks = sms * 8, iters = 4096;
    double* out;
    cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        k_dfma<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("DFMA: %.3f ms  %.2f TFLOP/s (2 flop per FMA)\n", ms, 2.0 * threads * blocks * (double)iters * 8 / ms * 1e-9);
        cudaEventRecord(e0);
        k_dmma<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        // one m8n8k4 per warp: 8*8*4 FMA
        if (rep) printf("DMMA m8n8k4: %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * (threads / 32) * blocks * (double)iters * 8 * 256 / ms * 1e-9);
    }
    printf("SMs %d\n", sms);
    return 0;
}
