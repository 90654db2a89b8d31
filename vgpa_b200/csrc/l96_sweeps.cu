// l96_sweeps.cu -- D = 40 forward and backward ODE sweeps, one CTA per inference
// problem, the whole sequential time loop on chip.
//
//   forward  : m' = -A m + b,  S' = -A S - S A^T + Sigma          (ode_solver.py:31-61)
//   backward : lam' = -dE/dm + A lam,  Psi' = -dE/dS + Psi A + A^T Psi  (:63-95)
//              with the observation jumps at obs_t (gaussian_like.py:200-243)
//              FUSED with the gradient assembly of VarGP.gradient
//              (variational.py:202-334): dL/dA[t], dL/db[t] are formed as soon
//              as lam[t], Psi[t] exist, so lam/Psi never travel to HBM.
//   solver tableaux exactly as src/numerics/{euler,heun,runge_kutta2,runge_kutta4}.py
//   (including runge_kutta2.py:96, where S stands in for A in the inner stage).
//
// CTA = 96 threads:
//   warps 0-1  "team": an 8 x 8 thread grid, each thread owns a 5 x 5 register tile
//              (rows ti+8r, cols tj+8c) of every 40 x 40 product; operands are read
//              from shared memory with conflict-free (pitch 42) broadcast loads and
//              the products run on the FP64 FMA pipe.
//   warp 2     "vector warp": the mean / lambda recurrences (40 x 40 mat-vecs), the
//              dL/db row, and ALL global->shared traffic, issued as 1-D bulk async
//              copies (TMA unit, UBLKCP) that complete on mbarriers one or two
//              stages ahead of their use.
// Symmetry: S and Psi are kept EXACTLY symmetric by forming P + P^T through a
// shared-memory transpose, so one product per RHS evaluation suffices.
#include "common.cuh"
#include "ptx.cuh"

namespace vgpa {
namespace {

constexpr int D = 40;
constexpr int P = 42;          // shared-memory row pitch (doubles): 336 B rows, 16 B aligned
constexpr int MAT = D * P;     // one padded matrix
constexpr int ROWB = D * 8;    // bytes of one matrix row in HBM
constexpr int TEAM = 64;
constexpr int NTH = 96;

enum { K_CUR = 0, K_MID = 1, K_NEXT = 2, K_SELF = 3 };

__host__ __device__ constexpr int n_stages(int m) { return m == ODE_EULER ? 1 : (m == ODE_RK4 ? 4 : 2); }
// which A (b / dE) a stage reads: the current index, the neighbour, their midpoint
__host__ __device__ constexpr int stage_kind(int m, int s)
{
    return m == ODE_EULER ? K_CUR
         : m == ODE_HEUN  ? (s == 0 ? K_CUR : K_NEXT)
         : m == ODE_RK2   ? (s == 0 ? K_CUR : K_MID)
                          : (s == 0 ? K_CUR : (s == 3 ? K_NEXT : K_MID));
}
// next stage operand = Y +/- next_coef * dt * k_s
__host__ __device__ constexpr double next_coef(int m, int s)
{
    return m == ODE_HEUN ? 1.0 : m == ODE_RK2 ? 0.5 : (s < 2 ? 0.5 : 1.0);
}
// weight of k_s in the final combination
__host__ __device__ constexpr double ksum_w(int m, int s)
{
    return m == ODE_RK2 ? (s == 0 ? 0.0 : 1.0) : m == ODE_RK4 ? ((s == 1 || s == 2) ? 2.0 : 1.0) : 1.0;
}
// Y_new = Y +/- final_step(ksum)
template <int METHOD>
__device__ __forceinline__ double final_step(double dt, double ksum)
{
    if (METHOD == ODE_HEUN) return (0.5 * dt) * ksum;
    if (METHOD == ODE_RK4) return dt * ksum / 6.0;
    return dt * ksum;
}

// ---- team product: acc[r][c] = sum_k L(ti+8r, k) * R(k, tj+8c) ----------------
// LK / RK: 0 = plain buffer, 1 = midpoint 0.5*(buf0 + buf1),
// LK = 2 : fused left operand isg[i]*L0[i][k] - 2*L1[i][k]   (gradient assembly)
template <int LK, int RK>
__device__ __forceinline__ void team_mm(const double* __restrict__ L0, const double* __restrict__ L1,
                                        const double* __restrict__ R0, const double* __restrict__ R1,
                                        const double* __restrict__ isg, int ti, int tj, double (&acc)[5][5])
{
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[r][c] = 0.0;
    double sc[5];
    if (LK == 2) {
#pragma unroll
        for (int r = 0; r < 5; ++r) sc[r] = isg[ti + 8 * r];
    }
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
        double a[5], b[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const int o = (ti + 8 * r) * P + k;
            if (LK == 0) a[r] = L0[o];
            else if (LK == 1) a[r] = 0.5 * (L0[o] + L1[o]);
            else a[r] = fma(sc[r], L0[o], -2.0 * L1[o]);
        }
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int o = k * P + tj + 8 * c;
            if (RK == 0) b[c] = R0[o];
            else b[c] = 0.5 * (R0[o] + R1[o]);
        }
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
}

__device__ __forceinline__ void tile_to_smem(double* __restrict__ T, int ti, int tj, const double (&acc)[5][5])
{
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) T[(ti + 8 * r) * P + tj + 8 * c] = acc[r][c];
}

// one row of  Aop v  for the vector warp
template <int KIND>
__device__ __forceinline__ double row_dot(const double* __restrict__ Ac, const double* __restrict__ An, int i,
                                          const double* __restrict__ v)
{
    double s = 0.0;
#pragma unroll 8
    for (int k = 0; k < D; ++k) {
        const int o = i * P + k;
        const double a = KIND == K_CUR ? Ac[o] : (KIND == K_NEXT ? An[o] : 0.5 * (Ac[o] + An[o]));
        s = fma(a, v[k], s);
    }
    return s;
}
template <int KIND>
__device__ __forceinline__ double pick(const double* __restrict__ c, const double* __restrict__ n, int i)
{
    return KIND == K_CUR ? c[i] : (KIND == K_NEXT ? n[i] : 0.5 * (c[i] + n[i]));
}

// issue the bulk copies of one 40 x 40 matrix (row by row into the padded tile)
__device__ __forceinline__ void load_matrix(double* dst, const double* src, uint64_t* bar, int lane)
{
    for (int i = lane; i < D; i += 32) bulk_g2s(dst + i * P, src + i * D, ROWB, bar);
}

// ===========================================================================
// forward sweep
// ===========================================================================
struct FwdSmem {
    double Sb[MAT], Hb[MAT], Tb[MAT], Ab[2][MAT];
    double bb[2][D], mv[D], vt[2][D], sig[D];
    uint64_t barA[2];
};

template <int METHOD>
__global__ void __launch_bounds__(NTH)
l96_fwd_kernel(Batch b, Scratch s, const double* __restrict__ x, long long xs, int p0)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
    constexpr int NS = n_stages(METHOD);
    const int tid = threadIdx.x, lane = tid & 31;
    const bool team = tid < TEAM;
    const int ti = tid >> 3, tj = tid & 7;
    const int lp = blockIdx.x, p = p0 + lp, N = b.N;
    const double* A = x + (long long)p * xs;
    const double* bo = A + (long long)N * D * D;
    double* mt = s.mt + (long long)lp * N * D;
    double* st = s.st + (long long)lp * N * D * D;
    const double dt = b.dt;

    if (tid == 0) {
        mbar_init(&sm.barA[0], 1);
        mbar_init(&sm.barA[1], 1);
        mbar_fence_init();
    }
    // initial state: S0 -> Sb (and trajectory slot 0), m0 -> mv
    for (int e = tid; e < D * D; e += NTH) {
        const int i = e / D, j = e % D;
        const double v = b.s0[p * b.s0_stride + e];
        sm.Sb[i * P + j] = v;
        st[e] = v;
    }
    if (tid < D) {
        const double v = b.m0[p * b.m0_stride + tid];
        sm.mv[tid] = v;
        mt[tid] = v;
        sm.sig[tid] = b.sigma[p * b.sigma_stride + tid];
    }
    __syncthreads();
    if (!team) {  // prologue loads: A_0, b_0 -> slot 0; A_1, b_1 -> slot 1
        for (int q = 0; q < 2 && q < N; ++q) {
            if (lane == 0) mbar_arrive_expect_tx(&sm.barA[q], D * ROWB + ROWB);
            __syncwarp();
            load_matrix(sm.Ab[q], A + (long long)q * D * D, &sm.barA[q], lane);
            if (lane == 0) bulk_g2s(sm.bb[q], bo + (long long)q * D, ROWB, &sm.barA[q]);
        }
    }
    uint32_t par[2] = {0u, 0u};
    mbar_wait(&sm.barA[0], par[0]);
    par[0] ^= 1u;

    for (int k = 0; k < N - 1; ++k) {
        const int cur = k & 1, nxt = cur ^ 1;
        const double* Ac = sm.Ab[cur];
        const double* An = sm.Ab[nxt];
        bool next_ready = false;
        double ksum[5][5];   // team: sum_s w_s k_s of the covariance tile
        double kv[2] = {0.0, 0.0};  // vector warp: same for its two rows of the mean
#pragma unroll
        for (int sidx = 0; sidx < NS; ++sidx) {
            // the covariance inner stage of RK2 uses S in place of A (runge_kutta2.py:96)
            const int kind = stage_kind(METHOD, sidx);
            const bool self = (METHOD == ODE_RK2 && sidx == 0);
            if (kind != K_CUR && !next_ready) {
                mbar_wait(&sm.barA[nxt], par[nxt]);
                par[nxt] ^= 1u;
                next_ready = true;
            }
            const double* X = (sidx == 0) ? sm.Sb : sm.Hb;
            double acc[5][5];
            if (team) {
                if (self)                 team_mm<0, 0>(X, nullptr, X, nullptr, nullptr, ti, tj, acc);
                else if (kind == K_CUR)   team_mm<0, 0>(Ac, nullptr, X, nullptr, nullptr, ti, tj, acc);
                else if (kind == K_NEXT)  team_mm<0, 0>(An, nullptr, X, nullptr, nullptr, ti, tj, acc);
                else                      team_mm<1, 0>(Ac, An, X, nullptr, nullptr, ti, tj, acc);
                tile_to_smem(sm.Tb, ti, tj, acc);
            } else {
                // mean stage: k = -Aop v + bop  on rows lane and lane+32
                const double* v = (sidx == 0) ? sm.mv : sm.vt[(sidx - 1) & 1];
                double ks[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int i = lane + 32 * q;
                    if (i < D) {
                        double av, bv;
                        if (kind == K_CUR)       { av = row_dot<K_CUR>(Ac, An, i, v);  bv = pick<K_CUR>(sm.bb[cur], sm.bb[nxt], i); }
                        else if (kind == K_NEXT) { av = row_dot<K_NEXT>(Ac, An, i, v); bv = pick<K_NEXT>(sm.bb[cur], sm.bb[nxt], i); }
                        else                     { av = row_dot<K_MID>(Ac, An, i, v);  bv = pick<K_MID>(sm.bb[cur], sm.bb[nxt], i); }
                        ks[q] = -av + bv;
                        const double w = ksum_w(METHOD, sidx);
                        if (w != 0.0) kv[q] = (sidx == 0 || (METHOD == ODE_RK2)) ? w * ks[q] : kv[q] + w * ks[q];
                        if (sidx < NS - 1)
                            sm.vt[sidx & 1][i] = sm.mv[i] + (next_coef(METHOD, sidx) * dt) * ks[q];
                    }
                }
            }
            __syncthreads();  // T complete (team) / next mean operand visible
            if (team) {
#pragma unroll
                for (int r = 0; r < 5; ++r)
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const int i = ti + 8 * r, j = tj + 8 * c;
                        const double kk = (i == j ? sm.sig[i] : 0.0) - (acc[r][c] + sm.Tb[j * P + i]);
                        const double w = ksum_w(METHOD, sidx);
                        if (w != 0.0) ksum[r][c] = (sidx == 0 || (METHOD == ODE_RK2)) ? w * kk : ksum[r][c] + w * kk;
                        if (sidx < NS - 1) {
                            sm.Hb[i * P + j] = sm.Sb[i * P + j] + (next_coef(METHOD, sidx) * dt) * kk;
                        } else {
                            const double sn = sm.Sb[i * P + j] + final_step<METHOD>(dt, ksum[r][c]);
                            sm.Sb[i * P + j] = sn;
                            st[(long long)(k + 1) * D * D + i * D + j] = sn;
                        }
                    }
            } else if (sidx == NS - 1) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int i = lane + 32 * q;
                    if (i < D) {
                        const double mn = sm.mv[i] + final_step<METHOD>(dt, kv[q]);
                        sm.mv[i] = mn;
                        mt[(long long)(k + 1) * D + i] = mn;
                    }
                }
            }
            __syncthreads();  // next operand (Hb / Sb) visible; T reusable
        }
        if (!next_ready && k + 1 < N) {  // Euler: A_{k+1} becomes "current" next step
            mbar_wait(&sm.barA[nxt], par[nxt]);
            par[nxt] ^= 1u;
        }
        // slot `cur` is dead: prefetch A_{k+2}, b_{k+2} into it
        if (!team && k + 2 < N) {
            if (lane == 0) mbar_arrive_expect_tx(&sm.barA[cur], D * ROWB + ROWB);
            __syncwarp();
            load_matrix(sm.Ab[cur], A + (long long)(k + 2) * D * D, &sm.barA[cur], lane);
            if (lane == 0) bulk_g2s(sm.bb[cur], bo + (long long)(k + 2) * D, ROWB, &sm.barA[cur]);
        }
    }
}

// ===========================================================================
// backward sweep + gradient assembly
// ===========================================================================
struct BwdSmem {
    double Pb[MAT], Hb[MAT], Tb[MAT], Ab[2][MAT], Gb[MAT], Sb[MAT];
    double gv[2][D], mv[D], bv[D], lam[D], lt[2][D], u[D], isg[D], Rv[D], jmv[D];
    uint64_t barA[2], barS, barG;
};

struct BwdArgs {
    const double* A;     // (N,D,D) of this launch's first problem (stride xs between problems)
    const double* bo;    // (N,D)   offsets (null when with_grad == 0)
    long long xs;
    const double* mt; const double* st;      // scratch (problem-major), may be null w/o grad
    const double* dEm; const double* dEs;    // (N,D), (N,D,D) per problem
    long long traj_v, traj_m;                // strides between problems of the above
    double* gA; double* gb; long long gs;    // gradient out (null: no gradient)
    const double* jm_dense; const double* js_dense;  // dense jump tables (stand-alone sweep) or null
    double* lam_out; double* psi_out;        // trajectories out (first problem only) or null
};

template <int METHOD>
__global__ void __launch_bounds__(NTH)
l96_bwd_kernel(Batch b, BwdArgs a, int p0)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
    constexpr int NS = n_stages(METHOD);
    const int tid = threadIdx.x, lane = tid & 31;
    const bool team = tid < TEAM;
    const int ti = tid >> 3, tj = tid & 7;
    const int lp = blockIdx.x, p = p0 + lp, N = b.N;
    const double* A = a.A + (long long)lp * a.xs;
    const double* bo = a.bo ? a.bo + (long long)lp * a.xs : nullptr;
    const double* mt = a.mt ? a.mt + (long long)lp * a.traj_v : nullptr;
    const double* st = a.st ? a.st + (long long)lp * a.traj_m : nullptr;
    const double* dEm = a.dEm + (long long)lp * a.traj_v;
    const double* dEs = a.dEs + (long long)lp * a.traj_m;
    const bool with_grad = a.gA != nullptr;
    double* gA = with_grad ? a.gA + (long long)lp * a.gs : nullptr;
    double* gb = with_grad ? a.gb + (long long)lp * a.gs : nullptr;
    const bool dense = a.jm_dense != nullptr;
    const bool keep = a.lam_out != nullptr && lp == 0;
    const double* oy = dense ? nullptr : b.obs_y + p * b.obs_y_stride;
    const double dt = b.dt, dtm = b.dt_model;
    const double theta = (b.theta != nullptr) ? b.theta[p * b.theta_stride] : 0.0;

    if (tid == 0) {
        mbar_init(&sm.barA[0], 1);
        mbar_init(&sm.barA[1], 1);
        mbar_init(&sm.barS, 1);
        mbar_init(&sm.barG, 1);
        mbar_fence_init();
    }
    for (int e = tid; e < MAT; e += NTH) sm.Pb[e] = 0.0;  // Psi[N-1] = 0
    if (tid < D) {
        sm.lam[tid] = 0.0;                                  // lam[N-1] = 0
        sm.isg[tid] = (b.sigma != nullptr) ? 1.0 / b.sigma[p * b.sigma_stride + tid] : 0.0;
        sm.Rv[tid] = (b.R != nullptr) ? b.R[p * b.R_stride + tid] : 1.0;
        sm.gv[(N - 1) & 1][tid] = dEm[(long long)(N - 1) * D + tid];
    }
    // dE/dS tile of the current index lives in registers
    double Greg[5][5];
    if (team) {
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c)
                Greg[r][c] = dEs[(long long)(N - 1) * D * D + (ti + 8 * r) * D + tj + 8 * c];
    }
    __syncthreads();
    if (!team) {
        const int t = N - 1;
        if (lane == 0) mbar_arrive_expect_tx(&sm.barA[t & 1], D * ROWB);
        __syncwarp();
        load_matrix(sm.Ab[t & 1], A + (long long)t * D * D, &sm.barA[t & 1], lane);
        if (t >= 1) {
            if (lane == 0) mbar_arrive_expect_tx(&sm.barA[(t - 1) & 1], D * ROWB);
            __syncwarp();
            load_matrix(sm.Ab[(t - 1) & 1], A + (long long)(t - 1) * D * D, &sm.barA[(t - 1) & 1], lane);
            if (lane == 0) mbar_arrive_expect_tx(&sm.barG, D * ROWB + ROWB);
            __syncwarp();
            load_matrix(sm.Gb, dEs + (long long)(t - 1) * D * D, &sm.barG, lane);
            if (lane == 0) bulk_g2s(sm.gv[(t - 1) & 1], dEm + (long long)(t - 1) * D, ROWB, &sm.barG);
        }
        if (with_grad) {
            if (lane == 0) mbar_arrive_expect_tx(&sm.barS, D * ROWB + 2 * ROWB);
            __syncwarp();
            load_matrix(sm.Sb, st + (long long)t * D * D, &sm.barS, lane);
            if (lane == 0) {
                bulk_g2s(sm.mv, mt + (long long)t * D, ROWB, &sm.barS);
                bulk_g2s(sm.bv, bo + (long long)t * D, ROWB, &sm.barS);
            }
        }
    }
    uint32_t parA[2] = {0u, 0u}, parS = 0u, parG = 0u;
    mbar_wait(&sm.barA[(N - 1) & 1], parA[(N - 1) & 1]);
    parA[(N - 1) & 1] ^= 1u;

    for (int t = N - 1; t >= 0; --t) {
        const int cur = t & 1, nxt = cur ^ 1;
        const double* Ac = sm.Ab[cur];
        const double* An = sm.Ab[nxt];
        if (keep) {  // lam[t], Psi[t] for vgpa_eval_full / the stand-alone sweep
            if (team) {
#pragma unroll
                for (int r = 0; r < 5; ++r)
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const int i = ti + 8 * r, j = tj + 8 * c;
                        a.psi_out[(long long)t * D * D + i * D + j] = sm.Pb[i * P + j];
                    }
            } else {
                for (int i = lane; i < D; i += 32) a.lam_out[(long long)t * D + i] = sm.lam[i];
            }
        }
        // ---- gradient at index t (variational.py:263-288) ----------------------
        if (with_grad) {
            mbar_wait(&sm.barS, parS);
            parS ^= 1u;
            double acc[5][5];
            if (team) {
                // W = (Sigma^-1 A_t - 2 Psi_t) S_t
                team_mm<2, 0>(Ac, sm.Pb, sm.Sb, nullptr, sm.isg, ti, tj, acc);
            } else {
                for (int i = lane; i < D; i += 32) {
                    const int f1 = (i + 1) % D, b1 = (i + D - 1) % D, b2 = (i + D - 2) % D;
                    // <f> of Lorenz 96 (lorenz_96.py:440-462)
                    const double Ef = (sm.Sb[f1 * P + b1] - sm.Sb[b2 * P + b1]) +
                                      (sm.mv[f1] - sm.mv[b2]) * sm.mv[b1] - sm.mv[i] + theta;
                    const double am = row_dot<K_CUR>(Ac, An, i, sm.mv);
                    const double db = sm.isg[i] * (-Ef - am + sm.bv[i]);  // variational.py:324-334
                    const double ui = db + sm.lam[i];
                    sm.u[i] = ui;
                    gb[(long long)t * D + i] = dtm * ui;                   // :280,285
                }
            }
            __syncthreads();
            if (team) {
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    const int i = ti + 8 * r;
                    const int f1 = (i + 1) % D, b1 = (i + D - 1) % D, b2 = (i + D - 2) % D;
                    // row i of <df/dx> S  (Jacobian of lorenz_96.py:34-83 applied to S)
                    const double cb1 = sm.mv[f1] - sm.mv[b2], cf = sm.mv[b1];
                    const double ui = sm.u[i], is = sm.isg[i];
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const int j = tj + 8 * c;
                        const double es = -sm.Sb[i * P + j] + cf * sm.Sb[f1 * P + j] - cf * sm.Sb[b2 * P + j] +
                                          cb1 * sm.Sb[b1 * P + j];
                        gA[(long long)t * D * D + i * D + j] = dtm * (acc[r][c] + is * es - ui * sm.mv[j]);
                    }
                }
            }
        }
        if (t == 0) break;
        // ---- one backward step t -> t-1 ------------------------------------------
        bool next_ready = false;
        double ksum[5][5];
        double kv[2] = {0.0, 0.0};
#pragma unroll
        for (int sidx = 0; sidx < NS; ++sidx) {
            const int kind = stage_kind(METHOD, sidx);
            if (kind != K_CUR && !next_ready) {
                mbar_wait(&sm.barA[nxt], parA[nxt]);
                parA[nxt] ^= 1u;
                mbar_wait(&sm.barG, parG);
                parG ^= 1u;
                next_ready = true;
            }
            const double* X = (sidx == 0) ? sm.Pb : sm.Hb;
            double acc[5][5];
            if (team) {
                // Q = X Aop
                if (kind == K_CUR)       team_mm<0, 0>(X, nullptr, Ac, nullptr, nullptr, ti, tj, acc);
                else if (kind == K_NEXT) team_mm<0, 0>(X, nullptr, An, nullptr, nullptr, ti, tj, acc);
                else                     team_mm<0, 1>(X, nullptr, Ac, An, nullptr, ti, tj, acc);
                tile_to_smem(sm.Tb, ti, tj, acc);
            } else {
                const double* v = (sidx == 0) ? sm.lam : sm.lt[(sidx - 1) & 1];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int i = lane + 32 * q;
                    if (i < D) {
                        double av, gk;
                        if (kind == K_CUR)       { av = row_dot<K_CUR>(Ac, An, i, v);  gk = pick<K_CUR>(sm.gv[cur], sm.gv[nxt], i); }
                        else if (kind == K_NEXT) { av = row_dot<K_NEXT>(Ac, An, i, v); gk = pick<K_NEXT>(sm.gv[cur], sm.gv[nxt], i); }
                        else                     { av = row_dot<K_MID>(Ac, An, i, v);  gk = pick<K_MID>(sm.gv[nxt], sm.gv[cur], i); }
                        const double ks = -gk + av;  // ode_solver.py:77
                        const double w = ksum_w(METHOD, sidx);
                        if (w != 0.0) kv[q] = (sidx == 0 || (METHOD == ODE_RK2)) ? w * ks : kv[q] + w * ks;
                        if (sidx < NS - 1) sm.lt[sidx & 1][i] = sm.lam[i] - (next_coef(METHOD, sidx) * dt) * ks;
                    }
                }
            }
            __syncthreads();  // T complete
            if (!team && sidx == 0 && with_grad) {
                // S_t, m_t, b_t are dead (gradient written): prefetch index t-1
                if (lane == 0) mbar_arrive_expect_tx(&sm.barS, D * ROWB + 2 * ROWB);
                __syncwarp();
                load_matrix(sm.Sb, st + (long long)(t - 1) * D * D, &sm.barS, lane);
                if (lane == 0) {
                    bulk_g2s(sm.mv, mt + (long long)(t - 1) * D, ROWB, &sm.barS);
                    bulk_g2s(sm.bv, bo + (long long)(t - 1) * D, ROWB, &sm.barS);
                }
            }
            if (team) {
#pragma unroll
                for (int r = 0; r < 5; ++r)
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const int i = ti + 8 * r, j = tj + 8 * c;
                        double g;
                        if (kind == K_CUR) g = Greg[r][c];
                        else if (kind == K_NEXT) g = sm.Gb[i * P + j];
                        else g = 0.5 * (sm.Gb[i * P + j] + Greg[r][c]);
                        const double kk = -g + (acc[r][c] + sm.Tb[j * P + i]);  // ode_solver.py:94
                        const double w = ksum_w(METHOD, sidx);
                        if (w != 0.0) ksum[r][c] = (sidx == 0 || (METHOD == ODE_RK2)) ? w * kk : ksum[r][c] + w * kk;
                        if (sidx < NS - 1) sm.Hb[i * P + j] = sm.Pb[i * P + j] - (next_coef(METHOD, sidx) * dt) * kk;
                    }
            }
            if (sidx < NS - 1) __syncthreads();  // Hb visible, T reusable
        }
        if (!next_ready) {  // Euler: index t-1 data becomes "current" next step
            mbar_wait(&sm.barA[nxt], parA[nxt]);
            parA[nxt] ^= 1u;
            mbar_wait(&sm.barG, parG);
            parG ^= 1u;
        }
        // ---- final combination + jump at index t-1 -----------------------------------
        const int n_obs = dense ? -1 : b.obs_index[t - 1];
        if (team) {
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const int i = ti + 8 * r, j = tj + 8 * c;
                    double pn = sm.Pb[i * P + j] - final_step<METHOD>(dt, ksum[r][c]);
                    if (dense) pn += a.js_dense[(long long)(t - 1) * D * D + i * D + j];
                    else if (n_obs >= 0 && i == j) pn += 0.5 / sm.Rv[i];  // gaussian_like.py:238
                    sm.Pb[i * P + j] = pn;
                    Greg[r][c] = sm.Gb[i * P + j];  // dE/dS[t-1] becomes current
                }
        } else {
            if (!dense && n_obs >= 0 && with_grad) mbar_wait(&sm.barS, parS);  // m[t-1] landed (parity unchanged)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = lane + 32 * q;
                if (i < D) {
                    double ln = sm.lam[i] - final_step<METHOD>(dt, kv[q]);
                    if (dense) ln += a.jm_dense[(long long)(t - 1) * D + i];
                    else if (n_obs >= 0) {
                        const double mprev = with_grad ? sm.mv[i] : mt[(long long)(t - 1) * D + i];
                        ln += -(oy[(long long)n_obs * D + i] - mprev) / sm.Rv[i];  // :235
                    }
                    sm.lam[i] = ln;
                }
            }
        }
        __syncthreads();  // Psi, lam of index t-1 complete; slot `cur`, Gb, gv[cur] dead
        if (!team && t >= 2) {
            if (lane == 0) mbar_arrive_expect_tx(&sm.barA[cur], D * ROWB);
            __syncwarp();
            load_matrix(sm.Ab[cur], A + (long long)(t - 2) * D * D, &sm.barA[cur], lane);
            if (lane == 0) mbar_arrive_expect_tx(&sm.barG, D * ROWB + ROWB);
            __syncwarp();
            load_matrix(sm.Gb, dEs + (long long)(t - 2) * D * D, &sm.barG, lane);
            if (lane == 0) bulk_g2s(sm.gv[cur], dEm + (long long)(t - 2) * D, ROWB, &sm.barG);
        }
    }
}

template <typename K>
void set_smem(K kernel, size_t bytes)
{
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

void launch_l96_fwd(const Batch& b, const Scratch& s, const double* x, long long xs, int p0, int count,
                    cudaStream_t st)
{
    const size_t sh = sizeof(FwdSmem);
    switch (b.method) {
    case ODE_EULER: set_smem(l96_fwd_kernel<ODE_EULER>, sh); l96_fwd_kernel<ODE_EULER><<<count, NTH, sh, st>>>(b, s, x, xs, p0); break;
    case ODE_HEUN:  set_smem(l96_fwd_kernel<ODE_HEUN>, sh);  l96_fwd_kernel<ODE_HEUN><<<count, NTH, sh, st>>>(b, s, x, xs, p0); break;
    case ODE_RK2:   set_smem(l96_fwd_kernel<ODE_RK2>, sh);   l96_fwd_kernel<ODE_RK2><<<count, NTH, sh, st>>>(b, s, x, xs, p0); break;
    default:        set_smem(l96_fwd_kernel<ODE_RK4>, sh);   l96_fwd_kernel<ODE_RK4><<<count, NTH, sh, st>>>(b, s, x, xs, p0); break;
    }
}

static void bwd_launch(const Batch& b, const BwdArgs& a, int p0, int count, cudaStream_t st)
{
    const size_t sh = sizeof(BwdSmem);
    switch (b.method) {
    case ODE_EULER: set_smem(l96_bwd_kernel<ODE_EULER>, sh); l96_bwd_kernel<ODE_EULER><<<count, NTH, sh, st>>>(b, a, p0); break;
    case ODE_HEUN:  set_smem(l96_bwd_kernel<ODE_HEUN>, sh);  l96_bwd_kernel<ODE_HEUN><<<count, NTH, sh, st>>>(b, a, p0); break;
    case ODE_RK2:   set_smem(l96_bwd_kernel<ODE_RK2>, sh);   l96_bwd_kernel<ODE_RK2><<<count, NTH, sh, st>>>(b, a, p0); break;
    default:        set_smem(l96_bwd_kernel<ODE_RK4>, sh);   l96_bwd_kernel<ODE_RK4><<<count, NTH, sh, st>>>(b, a, p0); break;
    }
}

void launch_l96_bwd(const Batch& b, const Scratch& s, const double* x, long long xs, double* grad,
                    long long gs, int p0, int count, const Extra& ex, cudaStream_t st)
{
    const long long N = b.N;
    BwdArgs a{};
    a.A = x + (long long)p0 * xs;
    a.bo = a.A + N * D * D;
    a.xs = xs;
    a.mt = s.mt; a.st = s.st; a.dEm = s.dEm; a.dEs = s.dEs;
    a.traj_v = N * D; a.traj_m = N * D * D;
    if (grad != nullptr) {
        a.gA = grad + (long long)p0 * gs;
        a.gb = a.gA + N * D * D;
        a.gs = gs;
    }
    a.lam_out = ex.lamt; a.psi_out = ex.psit;
    bwd_launch(b, a, p0, count, st);
}

void launch_bwd_dense_l96(int method, int N, double dt, const double* A, const double* dEm,
                          const double* dEs, const double* jm, const double* js, double* lam,
                          double* psi, cudaStream_t st)
{
    Batch b{};
    b.model = MODEL_L96; b.method = method; b.D = D; b.N = N; b.B = 1; b.dt = dt; b.dt_model = dt;
    BwdArgs a{};
    a.A = A; a.xs = 0; a.dEm = dEm; a.dEs = dEs;
    a.jm_dense = jm; a.js_dense = js; a.lam_out = lam; a.psi_out = psi;
    bwd_launch(b, a, 0, 1, st);
}

}  // namespace vgpa
