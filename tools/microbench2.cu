// microbench2.cu -- larger FP64 MMA shapes and DFMA/DMMA co-issue on sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int TILES>
__global__ void m16n8k4(double* out, int iters, double a, double b)
{
    double c[TILES][4];
    for (int i = 0; i < TILES; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < TILES; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(a));
    double s = 0; for (int i = 0; i < TILES; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456) out[0] = s;
}
template <int TILES>
__global__ void m16n8k8(double* out, int iters, double a, double b)
{
    double c[TILES][4];
    for (int i = 0; i < TILES; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < TILES; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    double s = 0; for (int i = 0; i < TILES; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456) out[0] = s;
}
template <int TILES>
__global__ void m16n8k16(double* out, int iters, double a, double b)
{
    double c[TILES][4];
    for (int i = 0; i < TILES; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < TILES; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    double s = 0; for (int i = 0; i < TILES; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456) out[0] = s;
}
// DMMA m8n8k4 and DFMA interleaved in one warp
template <int TILES, int NF>
__global__ void mixed(double* out, int iters, double a, double b)
{
    double c[TILES > 0 ? TILES : 1][2], f[NF > 0 ? NF : 1];
    for (int i = 0; i < TILES; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TILES; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0; for (int i = 0; i < TILES; ++i) s += c[i][0] + c[i][1];
    for (int i = 0; i < NF; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}
template <typename F> float time_ms(F launch)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms;
}
int main()
{
    double* out; CK(cudaMalloc(&out, 8));
    const int sms = 148, iters = 10000, th = 256, per = 2;
    const double warps = (double)(th / 32) * per * sms;
    float ms;
    ms = time_ms([&] { m16n8k4<8><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\": \"dmma_m16n8k4\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * 512 * 8 * iters * warps / ms * 1e-9);
    ms = time_ms([&] { m16n8k8<8><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\": \"dmma_m16n8k8\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * 1024 * 8 * iters * warps / ms * 1e-9);
    ms = time_ms([&] { m16n8k16<8><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\": \"dmma_m16n8k16\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * 2048 * 8 * iters * warps / ms * 1e-9);
    ms = time_ms([&] { mixed<8, 0><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\": \"mixed_dmma8_dfma0\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * (256 * 8) * iters * warps / ms * 1e-9);
    ms = time_ms([&] { mixed<8, 16><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\": \"mixed_dmma8_dfma16\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * (256 * 8 + 32 * 16) * iters * warps / ms * 1e-9);
    ms = time_ms([&] { mixed<8, 64><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\": \"mixed_dmma8_dfma64\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * (256 * 8 + 32 * 64) * iters * warps / ms * 1e-9);
    ms = time_ms([&] { mixed<0, 64><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\": \"mixed_dmma0_dfma64\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * (32 * 64) * iters * warps / ms * 1e-9);
    return 0;
}
