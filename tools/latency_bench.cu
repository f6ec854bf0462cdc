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

// FP64 tensor path: dependent chain of mma.sync.m8n8k4.f64 on one warp; independent ones (issue rate); and a DFMA chain
// on one warp while the three other warps of its sub-partition (warps 4, 8, 12 of a 512-thread CTA) issue independent
// tensor instructions back to back -- what a vector phase costs beside a tensor phase.
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(512) dmma_chain(double *out, long long *cyc, double a, double b, int n)
{
    const int tid = threadIdx.x, warp = tid >> 5;
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0, x = a + tid * 1e-9;
    long long t0, t1;
    if (warp == 0) {
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < n; ++i) dmma(c0, c1, a, b);
        t1 = clock64();
        if (tid == 0) cyc[0] = t1 - t0;
        t0 = clock64();
#pragma unroll 4
        for (int i = 0; i < n; i += 4) { dmma(c0, c1, a, b); dmma(e0, e1, a, b); dmma(g0, g1, a, b); dmma(h0, h1, a, b); }
        t1 = clock64();
        if (tid == 0) cyc[1] = t1 - t0;
    }
    __syncthreads();
    // phase 2: warp 0 runs a DFMA chain; warps 4, 8, 12 (same sub-partition) hammer the tensor path
    if (warp == 0) {
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < n; ++i) x = fma(x, b, a);
        t1 = clock64();
        if (tid == 0) cyc[2] = t1 - t0;
    } else if ((warp & 3) == 0) {
        for (int i = 0; i < 2 * n; i += 4) { dmma(c0, c1, a, b); dmma(e0, e1, a, b); dmma(g0, g1, a, b); dmma(h0, h1, a, b); }
    }
    __syncthreads();
    // phase 3: warp 0 runs a dependent tensor chain beside the same three warps
    if (warp == 0) {
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < n; ++i) dmma(c0, c1, a, b);
        t1 = clock64();
        if (tid == 0) cyc[3] = t1 - t0;
    } else if ((warp & 3) == 0) {
        for (int i = 0; i < 8 * n; i += 4) { dmma(c0, c1, a, b); dmma(e0, e1, a, b); dmma(g0, g1, a, b); dmma(h0, h1, a, b); }
    }
    out[tid] = c0 + c1 + e0 + e1 + g0 + g1 + h0 + h1 + x;
}

int main()
{
    {
        double *o2;
        long long *c2, h2[4];
        cudaMalloc(&o2, 8 * 512);
        cudaMalloc(&c2, 8 * 4);
        const int n2 = 4096;
        dmma_chain<<<1, 512>>>(o2, c2, 1.0000001, 0.9999999, n2);
        cudaMemcpy(h2, c2, sizeof(h2), cudaMemcpyDeviceToHost);
        printf("DMMA m8n8k4: dependent chain %.1f cycles each, four independent chains %.1f cycles each; DFMA chain beside three tensor warps on its sub-partition %.1f; dependent DMMA chain beside them %.1f\n",
               (double)h2[0] / n2, (double)h2[1] / n2, (double)h2[2] / n2, (double)h2[3] / n2);
    }
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
