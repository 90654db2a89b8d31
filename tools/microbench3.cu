// microbench3.cu -- does the FP64 pipe skip inactive half-warps?  DFMA chains with a lane
// mask: all 32 lanes, lanes 0-15, lanes 0-7, lane 0, even lanes (16 active, both halves).
// Same instruction count in every case; if "lanes 0-15" runs ~2x faster than "all", a
// predicated-off upper half-warp costs no pipe cycle.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void dfma_masked(double* out, int iters, double a, double b, unsigned mask)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    if ((mask >> (threadIdx.x & 31)) & 1u) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}
// dependent-chain latency of DFMA and of MUFU.RCP64H + Newton
__global__ void lat_dfma(double* out, int iters, double a, double b, long long* cyc)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) x = fma(x, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (x == 123.456) out[0] = x;
}
__global__ void lat_rcp(double* out, int iters, double a, long long* cyc)
{
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double r;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        x = r + a;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (x == 123.456) out[0] = x;
}
__global__ void lat_dmma(double* out, int iters, double a, double b, long long* cyc)
{
    double c0 = threadIdx.x, c1 = 1.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (c0 + c1 == 123.456) out[0] = c0;
}
__global__ void lat_shfl(double* out, int iters, long long* cyc)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (x == 123.456) out[0] = x;
}
__global__ void lat_lds(double* out, int iters, long long* cyc)
{
    __shared__ double buf[64];
    buf[threadIdx.x] = (threadIdx.x + 1) % 32;
    buf[threadIdx.x + 32] = 0;
    __syncthreads();
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) x = buf[(int)x];
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (x == 123.456) out[0] = x;
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
    long long* cyc; CK(cudaMallocManaged(&cyc, 8));
    const int sms = 148, iters = 20000, th = 256, per = 2;
    const unsigned masks[] = {0xffffffffu, 0x0000ffffu, 0x000000ffu, 0x00000001u, 0x55555555u, 0xffff0000u, 0x00ff00ffu};
    for (unsigned m : masks) {
        float ms = time_ms([&] { dfma_masked<16><<<sms * per, th>>>(out, iters, 1.0000001, 1e-9, m); });
        printf("{\"bench\": \"dfma_masked\", \"mask\": \"0x%08x\", \"ms\": %.3f, \"warp_inst_per_clk_per_smsp\": %.3f}\n", m, ms,
               16.0 * iters * (th / 32) * per / 4.0 / (ms * 1e-3 * 1.965e9));
    }
    const int n = 4096;
    lat_dfma<<<1, 32>>>(out, n, 1.0000001, 1e-9, cyc); CK(cudaDeviceSynchronize());
    printf("{\"bench\": \"latency\", \"op\": \"dfma\", \"cycles\": %.1f}\n", (double)cyc[0] / n);
    lat_rcp<<<1, 32>>>(out, n, 1.5, cyc); CK(cudaDeviceSynchronize());
    printf("{\"bench\": \"latency\", \"op\": \"rcp64h+dadd\", \"cycles\": %.1f}\n", (double)cyc[0] / n);
    lat_dmma<<<1, 32>>>(out, n, 1.0000001, 1e-9, cyc); CK(cudaDeviceSynchronize());
    printf("{\"bench\": \"latency\", \"op\": \"dmma_m8n8k4\", \"cycles\": %.1f}\n", (double)cyc[0] / n);
    lat_shfl<<<1, 32>>>(out, n, cyc); CK(cudaDeviceSynchronize());
    printf("{\"bench\": \"latency\", \"op\": \"shfl64+dadd\", \"cycles\": %.1f}\n", (double)cyc[0] / n);
    lat_lds<<<1, 32>>>(out, n, cyc); CK(cudaDeviceSynchronize());
    printf("{\"bench\": \"latency\", \"op\": \"lds64+cvt\", \"cycles\": %.1f}\n", (double)cyc[0] / n);
    return 0;
}
