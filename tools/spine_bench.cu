// spine_bench.cu -- unloaded timing of the factorisation spine of l96_energy.cu: ONE CTA, warp 0 runs
// spine_block for the five block columns (clock64 around each), warps 1-3 only answer the barriers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I vgpa_b200/csrc -o tools/dbg/spine_bench tools/spine_bench.cu
#include "../vgpa_b200/csrc/l96_energy.cu"
#include <cstdio>
#include <vector>
using namespace vgpa;
namespace vgpa { namespace {
__global__ void __launch_bounds__(128, 1) spine_kernel(const double* S, long long* out, double* Lout, int reps)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EnSmem& sm = *reinterpret_cast<EnSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int rep = 0; rep < reps; ++rep) {
        for (int e = tid; e < 1600; e += 128) sm.Cb[sm_idx(e / 40, e % 40)] = S[e];
        __syncthreads();
        if (warp == 0) {
            bool bad = false;
            double c[8];
            long long t[6];
            t[0] = clock64();
            spine_block<true>(sm, 0, lane, c, bad);
            t[1] = clock64();
#pragma unroll
            for (int kb = 1; kb < NB; ++kb) {
                spine_block<false>(sm, kb, lane, c, bad);
                t[kb + 1] = clock64();
            }
            if (lane == 0 && rep == reps - 1) {
                for (int i = 0; i < 5; ++i) out[i] = t[i + 1] - t[i];
                out[5] = bad;
            }
        } else {
            for (int kb = 0; kb < NB; ++kb) {
                bar_sync_par<BAR_L>(kb);
                if (kb < NB - 2) bar_arrive_par<BAR_T>(kb);
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < 1600; e += 128) Lout[e] = sm.Cb[sm_idx(e / 40, e % 40)];
}
} }
int main()
{
    const int Dd = 40;
    std::vector<double> G(Dd * Dd), S(Dd * Dd);
    srand(1);
    for (auto& v : G) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < Dd; ++i)
        for (int j = 0; j < Dd; ++j) {
            double a = (i == j) ? 1.0 : 0.0;
            for (int k = 0; k < Dd; ++k) a += G[i * Dd + k] * G[j * Dd + k];
            S[i * Dd + j] = a;
        }
    double *dS, *dL; long long* dout;
    cudaMalloc(&dS, 1600 * 8); cudaMalloc(&dL, 1600 * 8); cudaMalloc(&dout, 64);
    cudaMemcpy(dS, S.data(), 1600 * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(spine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EnSmem));
    spine_kernel<<<1, 128, sizeof(EnSmem)>>>(dS, dout, dL, 3);
    cudaError_t e = cudaDeviceSynchronize();
    long long out[6]; std::vector<double> L(1600);
    cudaMemcpy(out, dout, 48, cudaMemcpyDeviceToHost); cudaMemcpy(L.data(), dL, 1600 * 8, cudaMemcpyDeviceToHost);
    // host LDL^T for comparison of the unit-lower factor
    std::vector<double> C = S, d(Dd);
    double worst = 0.0;
    for (int j = 0; j < Dd; ++j) {
        d[j] = C[j * Dd + j];
        for (int i = j + 1; i < Dd; ++i) C[i * Dd + j] /= d[j];
        for (int i = j + 1; i < Dd; ++i)
            for (int m = j + 1; m <= i; ++m) C[i * Dd + m] -= C[i * Dd + j] * d[j] * C[m * Dd + j];
    }
    for (int i = 0; i < Dd; ++i)
        for (int j = 0; j < i; ++j) worst = fmax(worst, fabs(L[i * Dd + j] - C[i * Dd + j]));
    printf("{\"err\": \"%s\", \"block_cycles\": [%lld, %lld, %lld, %lld, %lld], \"bad\": %lld, \"max_abs_err_L\": %.3e}\n",
           cudaGetErrorString(e), out[0], out[1], out[2], out[3], out[4], out[5], worst);
    return 0;
}
