// latency_bench.cu -- dependent-chain latencies on one warp (cycles per operation): DFMA, DMUL, DADD, LDS, named barrier.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/latency_bench tools/latency_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chain(double *out, long long *cyc, double a, double b, int n)
{
    __shared__ int ism[1024];
    const int tid = threadIdx.x;
    for (int i = tid; i < 1024; i += blockDim.x) ism[i] = (i * 37 + 1) & 1023;
    __syncthreads();
    double x = a + tid * 1e-9;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fma(x, b, a);
    t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = x * b;
    t1 = clock64();
    if (tid == 0) cyc[1] = t1 - t0;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = x + b;
    t1 = clock64();
    if (tid == 0) cyc[2] = t1 - t0;
    // LDS pointer chase
    int idx = tid & 1023;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) idx = ism[idx];
    t1 = clock64();
    if (tid == 0) cyc[3] = t1 - t0;
    // division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < n; ++i) x = 1.0 / (x + 1.5);
    t1 = clock64();
    if (tid == 0) cyc[4] = t1 - t0;
    // __syncthreads
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) __syncthreads();
    t1 = clock64();
    if (tid == 0) cyc[5] = t1 - t0;
    // named barrier 96 threads (first three warps)
    if (blockDim.x >= 96 && tid < 96) {
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < n; ++i) asm volatile("bar.sync 1, 96;" ::: "memory");
        t1 = clock64();
        if (tid == 0) cyc[6] = t1 - t0;
    }
    // sincos (library)
    t0 = clock64();
    double s, c;
    for (int i = 0; i < n; ++i) { sincos(x, &s, &c); x = s + c; }
    t1 = clock64();
    if (tid == 0) cyc[7] = t1 - t0;
    // FFMA chain for comparison
    float f = (float)x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) f = fmaf(f, 1.0001f, 0.5f);
    t1 = clock64();
    if (tid == 0) cyc[8] = t1 - t0;
    out[tid] = x + idx + f;
}

int main()
{
    double *out;
    long long *cyc, h[9];
    cudaMalloc(&out, 8 * 1024);
    cudaMalloc(&cyc, 8 * 16);
    const int n = 4096;
    const char *names[9] = {"DFMA", "DMUL", "DADD", "LDS chase", "1/x (+add)", "__syncthreads", "bar.sync 96", "sincos+add", "FFMA"};
    for (int threads : {32, 128, 256}) {
        chain<<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, n);
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %d:", threads);
        for (int k = 0; k < 9; ++k) printf("  %s %.1f", names[k], (double)h[k] / n);
        printf("\n");
    }
    return 0;
}
