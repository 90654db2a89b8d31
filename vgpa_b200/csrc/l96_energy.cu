// l96_energy.cu -- time-parallel stage of the Lorenz-96 (D = 40) free energy:
// Esde(t), dEsde/dm(t), dEsde/dS(t) for every (problem, time index) pair, one CTA
// of 128 threads (4 warps) per pair.
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
//
// Blocked algorithm on 8 x 8 tiles (5 x 5 tile grid), bulk work on the FP64 tensor
// cores (mma.sync m8n8k4 f64, SASS DMMA):
//   for k = 0..4:  (a) 8 x 8 diagonal block L_kk by one warp, the whole lower triangle in the
//                      registers of every lane (no shuffles; LDL^T with hardware-seeded
//                      reciprocals, scaled at the end) -- the serial spine, so latency-tuned
//                  (b) panel L_ik by per-row triangular solves (i > k)
//                  (c) trailing C_ij -= L_ik L_jk^T as DMMA tiles, with look-ahead: the next
//                      diagonal block is factored while the other warps finish (c)
//   V = L^-1 afterwards: diagonal tiles by per-column solves, then block columns in parallel
//   (one warp per block column) as DMMA tile products
//   A L in place over A;  81 residual energies with warp-shuffle reductions;
//   V^T diag(d) V on the lower tiles, mirrored on store.
#include "common.cuh"
#include "ptx.cuh"

namespace vgpa {
namespace {

constexpr int D = 40;
constexpr int P = 44;               // pitch (doubles): DMMA fragment loads conflict-free
constexpr int MAT = D * P;
constexpr int ROWB = D * 8;
constexpr int K = 2 * D + 1;        // sigma points

constexpr int NTH = 128;
constexpr int NB = 5;               // 8 x 8 tile grid

struct EnSmem {
    double Cb[MAT];   // c S -> L (lower block triangle; upper tiles keep stale data, never read)
    double Wb[MAT];   // R -> V = L^-1 (lower block triangle, zero elsewhere)
    double Ab[MAT];   // A(t) -> A L
    double mv[D], bv[D], Am[D], isg[D], qv[D], dv[D];
    double dinv[D];   // 1 / L[j][j]
    double var[K + 3];
    double esde;
    int bad;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// L[i][col] with the upper triangle read as zero (upper tiles of Cb hold stale data)
__device__ __forceinline__ double l_at(const EnSmem& sm, int i, int col)
{
    const double v = sm.Cb[i * P + col];
    return (col <= i) ? v : 0.0;
}

// reciprocal and reciprocal square root from the hardware seed (MUFU.RCP64H / RSQ64H, ~20 bits)
// plus two Newton steps: full double precision without the long IEEE division / sqrt sequences,
// which sit on the serial critical path of the diagonal-block factorisation
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(fma(-hx * y, y, 0.5), y, y);
    y = fma(fma(-hx * y, y, 0.5), y, y);
    return y;
}

// ---- (a) diagonal 8 x 8 block: C_kk -> L_kk (into Cb), 1/L_jj (into dinv) -------------
// Latency is everything here (this is the serial spine of the factorisation), so every
// lane of the warp holds the WHOLE lower triangle (36 values) in registers and runs the
// LDL^T elimination redundantly: no shuffles, no shared-memory round trips; per pivot
// the dependent chain is reciprocal -> multiply -> one FMA.
__device__ __forceinline__ void factor_diag(EnSmem& sm, int k, int lane)
{
    double c[8][8], d[8];
    const double* src = &sm.Cb[(8 * k) * P + 8 * k];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) c[i][j] = src[i * P + j];   // warp-uniform (broadcast) loads
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double pj = c[j][j];
        bad |= !(pj > 0.0);
        d[j] = pj;
        const double rp = fast_rcp(pj);
#pragma unroll
        for (int i = j + 1; i < 8; ++i) {
            const double lij = c[i][j] * rp;           // unit-lower factor entry
#pragma unroll
            for (int m = j + 1; m <= i; ++m) c[i][m] = fma(-lij, c[m][j], c[i][m]);
        }
    }
    if (bad && lane == 0) sm.bad = 1;
    // scale: L = Lt D^1/2
    double rs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) rs[j] = fast_rsqrt(d[j]);
    if (lane == 0) {
        double* lo = &sm.Cb[(8 * k) * P + 8 * k];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                lo[i * P + j] = (j < i) ? c[i][j] * rs[j] : (j == i ? d[j] * rs[j] : 0.0);
#pragma unroll
        for (int j = 0; j < 8; ++j) sm.dinv[8 * k + j] = rs[j];
    }
}

// ---- (b) panel tile: solve X L_kk^T = C_ik; lane r < 8 owns row r of the 8 x 8 tile ----
__device__ __forceinline__ void panel_solve(EnSmem& sm, int i, int k, int lane)
{
    const int r = lane & 7;
    double x[8];
    double* row = &sm.Cb[(8 * i + r) * P + 8 * k];
    const double* Lk = &sm.Cb[(8 * k) * P + 8 * k];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = row[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double xj = x[j] * sm.dinv[8 * k + j];
        x[j] = xj;
#pragma unroll
        for (int m = j + 1; m < 8; ++m) x[m] = fma(-xj, Lk[m * P + j], x[m]);
    }
    if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) row[j] = x[j];
    }
}

// ---- T_kk = L_kk^-1 (into the diagonal tile of Wb); lane c < 8 owns column c ------------
__device__ __forceinline__ void invert_diag(EnSmem& sm, int k, int lane)
{
    const int c = lane & 7;
    const double* Lk = &sm.Cb[(8 * k) * P + 8 * k];
    double y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double acc = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int m = 0; m < i; ++m) acc = fma(-Lk[i * P + m], y[m], acc);
        y[i] = acc * sm.dinv[8 * k + i];
    }
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm.Wb[(8 * k + i) * P + 8 * k + c] = y[i];
    }
}

// one 8x8x8 tile product accumulate: acc += sum_{kk<8} Aop[m][kk] * Bop[kk][n]
// afun(kk) returns A[m=g][kk] for this lane's kk = h*4+q ; bfun(kk) returns B[kk][n=g]
#define TILE_MMA(acc0, acc1, AEXPR, BEXPR)                    \
    do {                                                      \
        _Pragma("unroll") for (int h = 0; h < 2; ++h) {       \
            const int kk = 4 * h + q;                         \
            const double a_ = (AEXPR);                        \
            const double b_ = (BEXPR);                        \
            dmma(acc0, acc1, a_, b_);                         \
        }                                                     \
    } while (0)

__global__ void __launch_bounds__(NTH, 5)
l96_energy_kernel(Batch b, Scratch s, const double* __restrict__ x, long long xs, int p0, int count, Extra ex)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EnSmem& sm = *reinterpret_cast<EnSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int N = b.N;
    const int lp = blockIdx.x / N, t = blockIdx.x - lp * N, p = p0 + lp;
    const double* At = x + (long long)p * xs + (long long)t * D * D;
    const double* bt = x + (long long)p * xs + (long long)N * D * D + (long long)t * D;
    const double* mt = s.mt + ((long long)lp * N + t) * D;
    const double* St = s.st + ((long long)lp * N + t) * D * D;
    const double theta = b.theta[p * b.theta_stride];
    const double kap = 1.05 * D, c = D + kap;                 // utilities.py:271
    const double w0 = kap / c, wi = 1.0 / (2.0 * c);          // :290-291

    if (tid == 0) sm.bad = 0;
    // S(t), A(t), m(t), b(t): 16-byte cp.async copies (LDGSTS) spread over the CTA
    cp_async_matrix(sm.Cb, St, P, tid, NTH);
    cp_async_matrix(sm.Ab, At, P, tid, NTH);
    cp_async_vector(sm.mv, mt, tid, 0);
    cp_async_vector(sm.bv, bt, tid, 32);
    cp_async_commit();
    if (tid < D) sm.isg[tid] = 1.0 / b.sigma[p * b.sigma_stride + tid];
    cp_async_wait<0>();
    __syncthreads();

    // <f>, <df/dx> for vgpa_eval_full (lorenz_96.py:34-83,440-462); S is still intact
    if (ex.Efx != nullptr && lp == 0) {
        for (int i = tid; i < D; i += NTH) {
            const int f1 = (i + 1) % D, b1 = (i + D - 1) % D, b2 = (i + D - 2) % D;
            ex.Efx[(long long)t * D + i] = (sm.Cb[f1 * P + b1] - sm.Cb[b2 * P + b1]) +
                                           (sm.mv[f1] - sm.mv[b2]) * sm.mv[b1] - sm.mv[i] + theta;
            double* row = ex.Edf + (long long)t * D * D + (long long)i * D;
            for (int j = 0; j < D; ++j) row[j] = 0.0;
            row[i] = -1.0;
            row[f1] = sm.mv[b1];
            row[b2] = -sm.mv[b1];
            row[b1] = sm.mv[f1] - sm.mv[b2];
        }
        __syncthreads();
    }
    // The factorisation runs on S itself: chol(c S) = sqrt(c) chol(S), so with L = chol(S),
    // V = L^-1 the sigma points are m +- sqrt(c) L[:, j] and the scale factors below
    // become sqrt(c)/2 and c/2 (numpy.linalg.cholesky reads the lower triangle; so do we).
    const double sqc = sqrt(c);
    // ---- blocked factorisation of S with the inverse carried along -----------------
    if (warp == 0) factor_diag(sm, 0, lane);
    __syncthreads();
    for (int k = 0; k < NB - 1; ++k) {
        // (b) panel rows i = k+1..4: one tile per warp
        if (warp < NB - 1 - k) panel_solve(sm, k + 1 + warp, k, lane);
        __syncthreads();
        // (c) trailing update C_ij -= L_ik L_jk^T with look-ahead: warp 0 updates the next
        //     diagonal tile and factors it at once, warps 1-3 share the other tiles
        if (warp == 0) {
            const int i = k + 1;
            double2 cc = *reinterpret_cast<double2*>(&sm.Cb[(8 * i + g) * P + 8 * i + 2 * q]);
            TILE_MMA(cc.x, cc.y, -sm.Cb[(8 * i + g) * P + 8 * k + kk], sm.Cb[(8 * i + g) * P + 8 * k + kk]);
            *reinterpret_cast<double2*>(&sm.Cb[(8 * i + g) * P + 8 * i + 2 * q]) = cc;
            __syncwarp();
            factor_diag(sm, k + 1, lane);
        } else {
            int n = 0;
            for (int i = k + 2; i < NB; ++i)
                for (int j = k + 1; j <= i; ++j) {
                    if ((n++ % 3) != warp - 1) continue;
                    double2 cc = *reinterpret_cast<double2*>(&sm.Cb[(8 * i + g) * P + 8 * j + 2 * q]);
                    TILE_MMA(cc.x, cc.y, -sm.Cb[(8 * i + g) * P + 8 * k + kk], sm.Cb[(8 * j + g) * P + 8 * k + kk]);
                    *reinterpret_cast<double2*>(&sm.Cb[(8 * i + g) * P + 8 * j + 2 * q]) = cc;
                }
        }
        __syncthreads();
    }

    // ---- V = L^-1, off the factorisation's critical path -------------------------------
    // diagonal tiles T_kk = L_kk^-1 (independent of each other) ...
    invert_diag(sm, warp, lane);
    if (warp == 0) invert_diag(sm, 4, lane);
    __syncthreads();
    // ... then block column j by warp j: V_ij = -T_ii sum_{m=j}^{i-1} L_im V_mj, i = j+1..4
    {
        const int j = warp;
        for (int i = j + 1; i < NB; ++i) {
            double s0 = 0.0, s1 = 0.0;
            for (int m = j; m < i; ++m)
                TILE_MMA(s0, s1, sm.Cb[(8 * i + g) * P + 8 * m + kk], sm.Wb[(8 * m + kk) * P + 8 * j + g]);
            *reinterpret_cast<double2*>(&sm.Wb[(8 * i + g) * P + 8 * j + 2 * q]) = make_double2(s0, s1);
            __syncwarp();
            double v0 = 0.0, v1 = 0.0;
            TILE_MMA(v0, v1, -sm.Wb[(8 * i + g) * P + 8 * i + kk], sm.Wb[(8 * i + kk) * P + 8 * j + g]);
            __syncwarp();
            *reinterpret_cast<double2*>(&sm.Wb[(8 * i + g) * P + 8 * j + 2 * q]) = make_double2(v0, v1);
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- A L in place over A (and A m): warp u owns tile-row u; tile-row 4 is shared ----
    {
        double a_own[NB][2], a_r4[NB][2];
#pragma unroll
        for (int Kb = 0; Kb < NB; ++Kb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                a_own[Kb][h] = sm.Ab[(8 * warp + g) * P + 8 * Kb + 4 * h + q];
                a_r4[Kb][h] = sm.Ab[(32 + g) * P + 8 * Kb + 4 * h + q];
            }
        double y = 0.0, y4 = 0.0;
#pragma unroll
        for (int Kb = 0; Kb < NB; ++Kb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double mk = sm.mv[8 * Kb + 4 * h + q];
                y = fma(a_own[Kb][h], mk, y);
                y4 = fma(a_r4[Kb][h], mk, y4);
            }
        y += __shfl_xor_sync(0xffffffffu, y, 1);
        y += __shfl_xor_sync(0xffffffffu, y, 2);
        y4 += __shfl_xor_sync(0xffffffffu, y4, 1);
        y4 += __shfl_xor_sync(0xffffffffu, y4, 2);
        if (q == 0) {
            sm.Am[8 * warp + g] = y;
            if (warp == 0) sm.Am[32 + g] = y4;
        }
        __syncthreads();   // every A fragment is in registers before any tile is overwritten
#pragma unroll
        for (int J = 0; J < NB; ++J) {   // own tile-row
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int Kb = J; Kb < NB; ++Kb)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    dmma(c0, c1, a_own[Kb][h], sm.Cb[(8 * Kb + 4 * h + q) * P + 8 * J + g]);
            *reinterpret_cast<double2*>(&sm.Ab[(8 * warp + g) * P + 8 * J + 2 * q]) = make_double2(c0, c1);
        }
#pragma unroll
        for (int J = 0; J < NB; ++J) {   // tile-row 4: J = 0,1,2 -> warps 0,1,2; J = 3,4 -> warp 3
            const int owner = J < 3 ? J : 3;
            if (owner != warp) continue;
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int Kb = J; Kb < NB; ++Kb)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    dmma(c0, c1, a_r4[Kb][h], sm.Cb[(8 * Kb + 4 * h + q) * P + 8 * J + g]);
            *reinterpret_cast<double2*>(&sm.Ab[(32 + g) * P + 8 * J + 2 * q]) = make_double2(c0, c1);
        }
    }
    __syncthreads();

    // ---- residual energies of the 81 sigma points: ONE THREAD PER SIGMA POINT walks the 40
    //      state entries with a sliding window (x[i-2], x[i-1], x[i], x[i+1]); lanes of a
    //      warp read consecutive columns of L and A L (conflict-free), the per-entry
    //      constants are warp-uniform broadcasts, and no cross-lane reduction is needed ----
    if (tid < K) {
        const int k = tid;
        const int kp = (k == 0) ? K - 1 : k - 1, kn = (k == K - 1) ? 0 : k + 1;
        const int col = (k == 0) ? 0 : ((k <= D) ? k - 1 : k - 1 - D);
        const int colp = (kp == 0) ? 0 : ((kp <= D) ? kp - 1 : kp - 1 - D);
        const int coln = (kn == 0) ? 0 : ((kn <= D) ? kn - 1 : kn - 1 - D);
        const double sg = (k == 0) ? 0.0 : ((k <= D) ? sqc : -sqc);     // +- sqrt(c): L is chol(S)
        const double sgp = (kp == 0) ? 0.0 : ((kp <= D) ? sqc : -sqc);
        const double sgn = (kn == 0) ? 0.0 : ((kn <= D) ? sqc : -sqc);
        // the flattened roll (lorenz_96.py:27-32) wraps into the neighbouring sigma points
        double xm2 = sm.mv[D - 2] + sgp * l_at(sm, D - 2, colp);
        double xm1 = sm.mv[D - 1] + sgp * l_at(sm, D - 1, colp);
        double x0 = sm.mv[0] + sg * l_at(sm, 0, col);
        const double xwrap = sm.mv[0] + sgn * l_at(sm, 0, coln);
        double var = 0.0;
#pragma unroll 8
        for (int i = 0; i < D; ++i) {
            const double xp1 = (i + 1 < D) ? sm.mv[i + 1] + sg * l_at(sm, i + 1, col) : xwrap;
            const double fx = (xp1 - xm2) * xm1 - x0 + theta;          // lorenz_96.py:85-101
            const double r = fx + ((sm.Am[i] - sm.bv[i]) + sg * sm.Ab[i * P + col]);
            var = fma(sm.isg[i], r * r, var);
            xm2 = xm1;
            xm1 = x0;
            x0 = xp1;
        }
        sm.var[k] = var;
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

    // ---- dEsde/dm = (c/2) Vc^T q = (sqrt(c)/2) V^T q  (Vc = V / sqrt(c)) ---------------------
    double* oEm = s.dEm + ((long long)lp * N + t) * D;
    double* oEs = s.dEs + ((long long)lp * N + t) * D * D;
    if (tid < D) {
        double a = 0.0;
        for (int k = tid; k < D; ++k) a = fma(sm.Wb[k * P + tid], sm.qv[k], a);
        oEm[tid] = 0.5 * sqc * a;
    }
    if (tid == 0) {
        s.esde_t[(long long)lp * N + t] = sm.esde;
        if (sm.bad) atomicCAS(&s.status[lp], 0, 1 + t);
    }
    // ---- dEsde/dS = (c^2/2) Vc^T diag(d) Vc = (c/2) V^T diag(d) V, lower tiles, mirrored ------
    {
        const double sc = 0.5 * c;
        // tile rows by cost (5 - I) blocks per tile: warp0: I=2, warp1: I=3, warp2: I=1, warp3: I=0 and 4
        for (int pass = 0; pass < 2; ++pass) {
            int I;
            if (pass == 0) I = (warp == 0) ? 2 : (warp == 1 ? 3 : (warp == 2 ? 1 : 0));
            else if (warp == 3) I = 4;
            else break;
            for (int J = 0; J <= I; ++J) {
                double c0 = 0.0, c1 = 0.0;
                for (int Kb = I; Kb < NB; ++Kb)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int kr = 8 * Kb + 4 * h + q;
                        dmma(c0, c1, sm.dv[kr] * sm.Wb[kr * P + 8 * I + g], sm.Wb[kr * P + 8 * J + g]);
                    }
                const int r = 8 * I + g, cc = 8 * J + 2 * q;
                const double v0 = sc * c0, v1 = sc * c1;
                if (I != J) {
                    *reinterpret_cast<double2*>(&oEs[(long long)r * D + cc]) = make_double2(v0, v1);
                    oEs[(long long)cc * D + r] = v0;
                    oEs[(long long)(cc + 1) * D + r] = v1;
                } else {   // diagonal tile: keep the lower triangle, mirror it
                    if (r >= cc) { oEs[(long long)r * D + cc] = v0; oEs[(long long)cc * D + r] = v0; }
                    if (r >= cc + 1) { oEs[(long long)r * D + cc + 1] = v1; oEs[(long long)(cc + 1) * D + r] = v1; }
                }
            }
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
