// microbench.cu -- FP64 pipe measurements that size the VGPA kernels' roofline.
//   1. DFMA peak (register-resident FMA chains) vs warps per SM and CTA shape
//   2. DMMA (mma.sync m8n8k4 f64) peak
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
// Prints one JSON object per line.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <int TILES>
__global__ void dmma_kernel(double* out, int iters, double a, double b)
{
    double c[TILES][2];
#pragma unroll
    for (int i = 0; i < TILES; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TILES; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TILES; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <typename F>
float time_ms(F launch)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    const int iters = 20000;
    // DFMA: CTA shapes (threads, ctas per SM)
    const int shapes[][2] = {{32, 1}, {64, 1}, {64, 2}, {64, 3}, {96, 2}, {96, 3}, {128, 1}, {128, 2}, {128, 4}, {256, 1}, {256, 2}, {256, 4}, {512, 2}, {1024, 1}, {1024, 2}};
    for (auto& sh : shapes) {
        const int th = sh[0], per = sh[1];
        float ms = time_ms([&] { dfma_kernel<16><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 16 * iters * (double)th * per * sms;
        printf("{\"bench\": \"dfma\", \"ilp\": 16, \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", th, per, ms, flops / ms * 1e-9);
    }
    for (auto& sh : shapes) {
        const int th = sh[0], per = sh[1];
        if (th > 256) continue;
        float ms = time_ms([&] { dfma_kernel<25><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 25 * iters * (double)th * per * sms;
        printf("{\"bench\": \"dfma\", \"ilp\": 25, \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", th, per, ms, flops / ms * 1e-9);
    }
    // DMMA m8n8k4: 8*8*4 = 256 FMA per warp-instruction
    const int mshapes[][2] = {{32, 1}, {64, 2}, {128, 1}, {128, 2}, {128, 4}, {256, 2}, {256, 4}, {512, 2}};
    for (auto& sh : mshapes) {
        const int th = sh[0], per = sh[1];
        float ms = time_ms([&] { dmma_kernel<8><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 256 * 8 * iters * (double)(th / 32) * per * sms;
        printf("{\"bench\": \"dmma_m8n8k4\", \"tiles\": 8, \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", th, per, ms, flops / ms * 1e-9);
    }
    CK(cudaFree(out));
    return 0;
}
