// fp64_peak.cu -- measure the FP64 roofline denominators on the box: DFMA (CUDA cores) and
// DMMA (mma.sync.m8n8k4.f64) throughput.  MEASURED_PEAKS.json has no FP64 entry, so the
// backward kernel's roofline fraction is quoted against the DFMA number printed here.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
    double acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    if (s == 12345.678) out[0] = s;
}

__global__ void dmma_kernel(double *out, int iters, double a, double b)
{
    double c0[4][2];
    for (int k = 0; k < 4; ++k) { c0[k][0] = threadIdx.x; c0[k][1] = k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[k][0]), "+d"(c0[k][1])
                         : "d"(a), "d"(b));
        }
    }
    double s = 0;
    for (int k = 0; k < 4; ++k) s += c0[k][0] + c0[k][1];
    if (s == 12345.678) out[0] = s;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double *out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int pass = 0; pass < 2; ++pass) {
            cudaEventRecord(e0);
            dfma_kernel<8><<<sms * (2048 / threads), threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (pass) {
                const double flops = 2.0 * 8 * iters * (double)threads * sms * (2048 / threads);
                printf("{\"kernel\": \"dfma\", \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", threads,
                       2048 / threads, ms, flops / ms * 1e-9);
            }
        }
    }
    for (int threads : {128, 512}) {
        // one CTA per SM (the backward kernel's residency) at 512 threads
        for (int pass = 0; pass < 2; ++pass) {
            cudaEventRecord(e0);
            dfma_kernel<8><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (pass) {
                const double flops = 2.0 * 8 * iters * (double)threads * sms;
                printf("{\"kernel\": \"dfma_1cta\", \"threads\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", threads, ms, flops / ms * 1e-9);
            }
        }
    }
    for (int threads : {256, 512, 1024}) {
        for (int pass = 0; pass < 2; ++pass) {
            cudaEventRecord(e0);
            dmma_kernel<<<sms * (2048 / threads), threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (pass) {
                const double flops = 2.0 * 8 * 8 * 4 * 4.0 * iters * (threads / 32.0) * sms * (2048 / threads);
                printf("{\"kernel\": \"dmma_m8n8k4\", \"threads\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", threads, ms, flops / ms * 1e-9);
            }
        }
    }
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    return 0;
}
