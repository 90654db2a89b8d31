// microbench4.cu -- single-warp issue cadence of FP64 instructions on sm_100a:
// cycles per DFMA for one resident warp per SM at ILP 1..16, and the same for DMMA.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
template <int ILP>
__global__ void dfma_ilp(double* out, int iters, double a, double b, long long* cyc)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (s == 123.456) out[0] = s;
}
template <int ILP>
__global__ void dmma_ilp(double* out, int iters, double a, double b, long long* cyc)
{
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (s == 123.456) out[0] = s;
}
template <int ILP> void run(double* out, long long* cyc, int warps)
{
    const int iters = 2000;
    dfma_ilp<ILP><<<148, 32 * warps>>>(out, iters, 1.0000001, 1e-9, cyc); CK(cudaDeviceSynchronize());
    printf("{\"bench\": \"dfma_cadence\", \"warps_per_sm\": %d, \"ilp\": %d, \"cycles_per_inst_per_warp\": %.2f}\n", warps, ILP, (double)cyc[0] / (iters * ILP));
    dmma_ilp<ILP><<<148, 32 * warps>>>(out, iters, 1.0000001, 1e-9, cyc); CK(cudaDeviceSynchronize());
    printf("{\"bench\": \"dmma_cadence\", \"warps_per_sm\": %d, \"ilp\": %d, \"cycles_per_inst_per_warp\": %.2f}\n", warps, ILP, (double)cyc[0] / (iters * ILP));
}
int main()
{
    double* out; CK(cudaMalloc(&out, 8));
    long long* cyc; CK(cudaMallocManaged(&cyc, 8));
    for (int warps : {1, 4, 8}) {
        run<1>(out, cyc, warps); run<2>(out, cyc, warps); run<4>(out, cyc, warps); run<8>(out, cyc, warps); run<16>(out, cyc, warps);
    }
    return 0;
}
