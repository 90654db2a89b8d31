// l96_energy.cu -- time-parallel stage of the Lorenz-96 (D = 40) free energy:
// Esde(t), dEsde/dm(t), dEsde/dS(t) for every (problem, time index) pair, one CTA
// of 128 threads per pair.
//
// The reference evaluates these with the unscented transform over 2D+1 = 81 sigma
// points (lorenz_96.py:389-418, utilities.py:239-310, variational.py:339-400),
// spending 81 dense solves per time index.  Here the same quantities come from ONE
// factorisation.  With c = D + kappa = 2.05 D, L = chol(c S) (lower), V = L^-1:
//   chi_0 = m, chi_{+-j} = m +- L[:, j]                          (utilities.py:283-288)
//   r_k   = l96(chi)_k + A chi_k - b = f_k + A m - b +- (A L)[:, j]
//   var_k = sum_i r_{k,i}^2 / sigma_i ,  Esde = 1/2 sum_k w_k var_k    (lorenz_96.py:398-401)
//   S^-1 (chi_{+-j} - m) = +- c V^T[:, j]  and  S^-1 = c V^T V,  hence
//   dEsde/dm = (c/2)  V^T q ,                q_j = w_i (var_{+j} - var_{-j})
//   dEsde/dS = (c^2/2) V^T diag(d) V ,       d_j = w_i (var_{+j} + var_{-j}) / 2 - Esde / c
// which is algebraically the reference's  dmS[:D] - Esde S^-1 m  and
// 0.5 (dmS[D:] - Esde S^-1)  (lorenz_96.py:414-418).  The discarded y_cov product of
// ut_approx (utilities.py:302-306) is not formed.
// l96() on the 81 x 40 sigma-point matrix uses numba's FLATTENED np.roll
// (lorenz_96.py:27-32,85-101): neighbours wrap across adjacent sigma points.
#include "common.cuh"
#include "ptx.cuh"

namespace vgpa {
namespace {

constexpr int D = 40;
constexpr int P = 42;
constexpr int MAT = D * P;
constexpr int ROWB = D * 8;
constexpr int K = 2 * D + 1;        // sigma points
constexpr int TOT = K * D;          // flattened sigma-point matrix
constexpr int NTH = 128;
constexpr int NPOS = D * (D + 1) / 2;             // lower-triangle positions
constexpr int PER = (NPOS + NTH - 1) / NTH;       // positions owned per thread (7)

struct EnSmem {
    double Sb[MAT];   // S(t); later A L
    double Lb[MAT];   // L = chol(c S), upper triangle zero
    double Wb[MAT];   // V = L^-1,       upper triangle zero
    double Ab[MAT];   // A(t)
    double col[2][D], wrow[2][D];   // pivot panels of the factorisation
    double mv[D], bv[D], Am[D], isg[D], dvec[D], rs[D], qv[D], dv[D];
    double var[K + 3];
    double esde;
    uint64_t bar;
    int bad;
};

// sigma point matrix entry at flattened index q (row k = q / D, column i = q % D)
__device__ __forceinline__ double chi_at(const EnSmem& sm, int q)
{
    const int k = q / D, i = q - k * D;
    const double m = sm.mv[i];
    if (k == 0) return m;
    if (k <= D) return m + sm.Lb[i * P + (k - 1)];
    return m - sm.Lb[i * P + (k - 1 - D)];
}

__global__ void __launch_bounds__(NTH)
l96_energy_kernel(Batch b, Scratch s, const double* __restrict__ x, long long xs, int p0, int count, Extra ex)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EnSmem& sm = *reinterpret_cast<EnSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = b.N;
    const int lp = blockIdx.x / N, t = blockIdx.x - lp * N, p = p0 + lp;
    const double* At = x + (long long)p * xs + (long long)t * D * D;
    const double* bt = x + (long long)p * xs + (long long)N * D * D + (long long)t * D;
    const double* mt = s.mt + ((long long)lp * N + t) * D;
    const double* St = s.st + ((long long)lp * N + t) * D * D;
    const double theta = b.theta[p * b.theta_stride];
    const double kap = 1.05 * D, c = D + kap;                 // utilities.py:271
    const double w0 = kap / c, wi = 1.0 / (2.0 * c);          // :290-291

    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
        sm.bad = 0;
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) mbar_arrive_expect_tx(&sm.bar, 2 * D * ROWB + 2 * ROWB);
        __syncwarp();
        for (int i = lane; i < D; i += 32) {
            bulk_g2s(sm.Sb + i * P, St + i * D, ROWB, &sm.bar);
            bulk_g2s(sm.Ab + i * P, At + i * D, ROWB, &sm.bar);
        }
        if (lane == 0) {
            bulk_g2s(sm.mv, mt, ROWB, &sm.bar);
            bulk_g2s(sm.bv, bt, ROWB, &sm.bar);
        }
    }
    if (tid < D) sm.isg[tid] = 1.0 / b.sigma[p * b.sigma_stride + tid];
    // decode the lower-triangle positions this thread owns: e = i (i + 1) / 2 + k
    int pi[PER], pk[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int e = tid + q * NTH;
        int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        while (i * (i + 1) / 2 > e) --i;
        pi[q] = (e < NPOS) ? i : -1;
        pk[q] = e - i * (i + 1) / 2;
    }
    mbar_wait(&sm.bar, 0u);

    // <f>, <df/dx> for vgpa_eval_full (lorenz_96.py:34-83,440-462); S is still intact
    if (ex.Efx != nullptr && lp == 0) {
        for (int i = tid; i < D; i += NTH) {
            const int f1 = (i + 1) % D, b1 = (i + D - 1) % D, b2 = (i + D - 2) % D;
            ex.Efx[(long long)t * D + i] = (sm.Sb[f1 * P + b1] - sm.Sb[b2 * P + b1]) +
                                           (sm.mv[f1] - sm.mv[b2]) * sm.mv[b1] - sm.mv[i] + theta;
            double* row = ex.Edf + (long long)t * D * D + (long long)i * D;
            for (int j = 0; j < D; ++j) row[j] = 0.0;
            row[i] = -1.0;
            row[f1] = sm.mv[b1];
            row[b2] = -sm.mv[b1];
            row[b1] = sm.mv[f1] - sm.mv[b2];
        }
    }

    // ---- factorisation of c S with the inverse carried along -------------------
    // Each thread keeps its positions of C = c S (-> unscaled factor) and of
    // W (-> unit-lower inverse) in registers; per pivot only column j of C and
    // row j of W go through shared memory.
    double cv[PER], wv[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        cv[q] = 0.0;
        wv[q] = 0.0;
        if (pi[q] >= 0) {
            cv[q] = c * sm.Sb[pi[q] * P + pk[q]];   // numpy.linalg.cholesky reads the lower triangle
            wv[q] = (pi[q] == pk[q]) ? 1.0 : 0.0;
            if (pk[q] == 0) sm.col[0][pi[q]] = cv[q];
            if (pi[q] == 0) sm.wrow[0][0] = 1.0;
        }
    }
    __syncthreads();
    for (int j = 0; j < D; ++j) {
        const int buf = j & 1;
        const double piv = sm.col[buf][j];
        if (!(piv > 0.0)) {
            if (tid == 0) sm.bad = 1;
        }
        const double rp = 1.0 / piv;
        if (tid == 0) sm.dvec[j] = piv;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int i = pi[q], k = pk[q];
            if (i > j) {
                const double li = sm.col[buf][i] * rp;
                if (k > j) cv[q] = fma(-li, sm.col[buf][k], cv[q]);
                else       wv[q] = fma(-li, sm.wrow[buf][k], wv[q]);
                if (k == j + 1) sm.col[buf ^ 1][i] = cv[q];
                if (i == j + 1) sm.wrow[buf ^ 1][k] = wv[q];
            }
        }
        __syncthreads();
    }
    if (tid < D) sm.rs[tid] = 1.0 / sqrt(sm.dvec[tid]);
    for (int e = tid; e < MAT; e += NTH) {
        sm.Lb[e] = 0.0;
        sm.Wb[e] = 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        if (pi[q] >= 0) {
            sm.Lb[pi[q] * P + pk[q]] = cv[q] * sm.rs[pk[q]];   // L[i][k] = C[i][k] / sqrt(D_k)
            sm.Wb[pi[q] * P + pk[q]] = wv[q] * sm.rs[pi[q]];   // V[i][k] = W[i][k] / sqrt(D_i)
        }
    }
    // A m
    if (tid < D) {
        double a = 0.0;
        for (int k = 0; k < D; ++k) a = fma(sm.Ab[tid * P + k], sm.mv[k], a);
        sm.Am[tid] = a;
    }
    __syncthreads();

    // ---- A L into the S buffer ----------------------------------------------------------
    const int ti = tid >> 4, tj = tid & 15;
    const int j2 = (tj + 32 < D) ? tj + 32 : D - 1;
    {
        double acc[5][3];
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) acc[r][cc] = 0.0;
#pragma unroll 4
        for (int k = 0; k < D; ++k) {
            double a[5], bb[3];
#pragma unroll
            for (int r = 0; r < 5; ++r) a[r] = sm.Ab[(ti + 8 * r) * P + k];
            bb[0] = sm.Lb[k * P + tj];
            bb[1] = sm.Lb[k * P + tj + 16];
            bb[2] = sm.Lb[k * P + j2];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) acc[r][cc] = fma(a[r], bb[cc], acc[r][cc]);
        }
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            sm.Sb[(ti + 8 * r) * P + tj] = acc[r][0];
            sm.Sb[(ti + 8 * r) * P + tj + 16] = acc[r][1];
            if (tj + 32 < D) sm.Sb[(ti + 8 * r) * P + tj + 32] = acc[r][2];
        }
    }
    __syncthreads();

    // ---- residual energies of the 81 sigma points (one warp per sigma point) ----
    for (int k = warp; k < K; k += NTH / 32) {
        const int colj = (k == 0) ? 0 : ((k <= D) ? k - 1 : k - 1 - D);
        const double sg = (k == 0) ? 0.0 : ((k <= D) ? 1.0 : -1.0);
        double part = 0.0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;
            if (i < D) {
                const int q = k * D + i;
                const int q1 = (q + 1 == TOT) ? 0 : q + 1;
                const int qm1 = (q == 0) ? TOT - 1 : q - 1;
                const int qm2 = (q < 2) ? q - 2 + TOT : q - 2;
                const double f = (chi_at(sm, q1) - chi_at(sm, qm2)) * chi_at(sm, qm1) - chi_at(sm, q) + theta;
                const double r = f + (sm.Am[i] + sg * sm.Sb[i * P + colj]) - sm.bv[i];
                part += sm.isg[i] * (r * r);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) sm.var[k] = part;
    }
    __syncthreads();
    if (warp == 0) {
        // Esde(t) = 1/2 sum_k w_k var_k  (fixed order: lanes stride the 81 values)
        double e = 0.0;
        for (int k = lane; k < K; k += 32) e += (k == 0 ? w0 : wi) * sm.var[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        e *= 0.5;
        if (lane == 0) sm.esde = e;
        for (int j = lane; j < D; j += 32) {
            const double vp = sm.var[1 + j], vm = sm.var[1 + D + j];
            sm.qv[j] = wi * (vp - vm);
            sm.dv[j] = 0.5 * (wi * (vp + vm)) - e / c;
        }
    }
    __syncthreads();

    // ---- dEsde/dm = (c/2) V^T q ------------------------------------------------------------
    double* oEm = s.dEm + ((long long)lp * N + t) * D;
    double* oEs = s.dEs + ((long long)lp * N + t) * D * D;
    if (tid < D) {
        double a = 0.0;
        for (int k = tid; k < D; ++k) a = fma(sm.Wb[k * P + tid], sm.qv[k], a);
        oEm[tid] = 0.5 * c * a;
    }
    if (tid == 0) {
        s.esde_t[(long long)lp * N + t] = sm.esde;
        if (sm.bad) atomicCAS(&s.status[lp], 0, 1 + t);
    }
    // ---- dEsde/dS = (c^2/2) V^T diag(d) V -----------------------------------------------
    {
        double acc[5][3];
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) acc[r][cc] = 0.0;
#pragma unroll 4
        for (int k = 0; k < D; ++k) {
            const double dk = sm.dv[k];
            double a[5], bb[3];
#pragma unroll
            for (int r = 0; r < 5; ++r) a[r] = sm.Wb[k * P + ti + 8 * r];
            bb[0] = dk * sm.Wb[k * P + tj];
            bb[1] = dk * sm.Wb[k * P + tj + 16];
            bb[2] = dk * sm.Wb[k * P + j2];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) acc[r][cc] = fma(a[r], bb[cc], acc[r][cc]);
        }
        const double sc = 0.5 * c * c;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            double* row = oEs + (long long)(ti + 8 * r) * D;
            row[tj] = sc * acc[r][0];
            row[tj + 16] = sc * acc[r][1];
            if (tj + 32 < D) row[tj + 32] = sc * acc[r][2];
        }
    }
}

}  // namespace

void launch_l96_energy(const Batch& b, const Scratch& s, const double* x, long long xs, int p0, int count,
                       const Extra& ex, cudaStream_t st)
{
    const size_t sh = sizeof(EnSmem);
    cudaFuncSetAttribute(l96_energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    const unsigned grid = (unsigned)((long long)count * b.N);
    l96_energy_kernel<<<grid, NTH, sh, st>>>(b, s, x, xs, p0, count, ex);
}

}  // namespace vgpa
