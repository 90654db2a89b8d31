// energy_prof.cu -- stand-alone timing harness for l96_energy_kernel: random SPD S(t), random
// A(t), one launch of `problems` x N items; prints the kernel time and (VGPA_EN_PROF) the
// average clock64 cycles CTA thread 0 spends in each phase.  Development aid.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I vgpa_b200/csrc -o tools/energy_prof tools/energy_prof.cu
#ifndef VGPA_EN_NOPROF
#define VGPA_EN_PROF
#endif
#include <cstdlib>
#ifndef VGPA_EN_SRC
#define VGPA_EN_SRC "../vgpa_b200/csrc/l96_energy.cu"
#endif
#include VGPA_EN_SRC
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
using namespace vgpa;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
int main(int argc, char** argv)
{
    const int problems = argc > 1 ? atoi(argv[1]) : 148, N = argc > 2 ? atoi(argv[2]) : 101, Dd = 40;
    const long long items = (long long)problems * N;
    const long long nx = (long long)N * Dd * (Dd + 1);
    std::vector<double> hx(nx), hS((size_t)N * Dd * Dd), hm((size_t)N * Dd);
    srand(1);
    auto rnd = [] { return rand() / (double)RAND_MAX - 0.5; };
    for (auto& v : hx) v = rnd();
    for (int t = 0; t < N; ++t) {   // S = G G^T + I
        std::vector<double> G(Dd * Dd);
        for (auto& v : G) v = rnd();
        for (int i = 0; i < Dd; ++i)
            for (int j = 0; j < Dd; ++j) {
                double a = (i == j) ? 1.0 : 0.0;
                for (int k = 0; k < Dd; ++k) a += G[i * Dd + k] * G[j * Dd + k];
                hS[((size_t)t * Dd + i) * Dd + j] = a;
            }
        for (int i = 0; i < Dd; ++i) hm[(size_t)t * Dd + i] = rnd();
    }
    double *x, *mt, *st, *dEm, *dEs, *es, *theta, *sigma; int* status;
    CK(cudaMalloc(&x, nx * 8)); CK(cudaMemcpy(x, hx.data(), nx * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&mt, items * Dd * 8)); CK(cudaMalloc(&st, items * Dd * Dd * 8));
    CK(cudaMalloc(&dEm, items * Dd * 8)); CK(cudaMalloc(&dEs, items * Dd * Dd * 8));
    CK(cudaMalloc(&es, items * 8)); CK(cudaMalloc(&status, problems * 4)); CK(cudaMemset(status, 0, problems * 4));
    for (int p = 0; p < problems; ++p) {
        CK(cudaMemcpy(mt + (long long)p * N * Dd, hm.data(), (size_t)N * Dd * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(st + (long long)p * N * Dd * Dd, hS.data(), (size_t)N * Dd * Dd * 8, cudaMemcpyHostToDevice));
    }
    std::vector<double> sg(Dd, 4.0); double th = 8.0;
    CK(cudaMalloc(&theta, 8)); CK(cudaMemcpy(theta, &th, 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&sigma, Dd * 8)); CK(cudaMemcpy(sigma, sg.data(), Dd * 8, cudaMemcpyHostToDevice));
    Batch b{}; b.model = MODEL_L96; b.D = Dd; b.N = N; b.B = problems; b.theta = theta; b.theta_stride = 0; b.sigma = sigma; b.sigma_stride = 0;
    Scratch s{}; s.mt = mt; s.st = st; s.dEm = dEm; s.dEs = dEs; s.esde_t = es; s.status = status;
    Extra ex{};
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 3; ++rep) {
        unsigned long long zero[32] = {0};
#ifdef VGPA_EN_PROF
        CK(cudaMemcpyToSymbol(g_prof, zero, sizeof(zero)));
#endif
        CK(cudaEventRecord(e0));
        launch_l96_energy(b, s, x, 0, 0, problems, ex, nullptr);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long pr[32] = {0};
#ifdef VGPA_EN_PROF
        CK(cudaMemcpyFromSymbol(pr, g_prof, sizeof(pr)));
#endif
        printf("{\"items\": %lld, \"ms\": %.3f, \"ns_per_item\": %.2f, \"phase_cycles\": [", items, ms, ms * 1e6 / items);
        double tot = 0;
        for (int i = 0; i < 24; ++i) { printf("%s%.0f", i ? ", " : "", (double)pr[i] / items); if (i < 12) tot += (double)pr[i] / items; }
        std::vector<double> hes(8); CK(cudaMemcpy(hes.data(), es, 64, cudaMemcpyDeviceToHost));
        int hst; CK(cudaMemcpy(&hst, status, 4, cudaMemcpyDeviceToHost));
        printf("], \"total_cycles\": %.0f, \"esde0\": %.12g, \"status0\": %d}\n", tot, hes[0], hst);
    }
    if (argc > 3) {   // dump the outputs of the first problem (N items) for an A/B comparison of kernel versions
        std::vector<double> o((size_t)N * (1 + Dd + Dd * Dd));
        CK(cudaMemcpy(o.data(), es, (size_t)N * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(o.data() + N, dEm, (size_t)N * Dd * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(o.data() + N + (size_t)N * Dd, dEs, (size_t)N * Dd * Dd * 8, cudaMemcpyDeviceToHost));
        FILE* f = fopen(argv[3], "wb");
        fwrite(o.data(), 8, o.size(), f);
        fclose(f);
    }
    return 0;
}
