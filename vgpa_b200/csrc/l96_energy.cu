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
// Blocked algorithm on 8 x 8 tiles (5 x 5 tile grid).  Scalar FP64 instructions and DMMA
// share ONE pipe on sm_100a (a warp-wide DFMA costs 2.26 pipe cycles whatever the lane
// mask, a DMMA 16: profiles/microbench_r02.jsonl), so everything that can be a tile
// product is one, and all tile loops are unrolled at compile time (immediate offsets):
//   load : the lower block triangle of S(t) and A(t), m(t), b(t) by 1-D bulk async copies
//          (TMA unit, SASS UBLKCP) on one mbarrier; the upper tiles of the L and V buffers
//          are zero-filled while the copies are in flight
//   for k = 0..4:  (a) diagonal block by one warp, the whole lower triangle in the registers
//                      of every lane: L_kk (LDL^T with hardware-seeded reciprocals) AND its
//                      inverse T_kk = L_kk^-1 (lane c solves column c) -- the serial spine
//                  (b) panel L_ik = C_ik T_kk^T as DMMA tile products (i > k)
//                  (c) trailing C_ij -= L_ik L_jk^T as DMMA tiles, with look-ahead: the next
//                      diagonal block is factored while the other warps finish (c)
//   V = L^-1: diagonal tiles are the T_kk; block column j by warp j as DMMA tile products
//   A L in place over A (and A m);  81 residual energies, one thread per sigma point;
//   V^T diag(d) V on the lower tiles, mirrored on store.
#include "common.cuh"
#include "ptx.cuh"

namespace vgpa {
namespace {

// optional phase timing (tools/energy_prof.cu defines VGPA_EN_PROF and includes this file)
#ifdef VGPA_EN_PROF
__device__ unsigned long long g_prof[32];
#define PROF_MARK(i)                                                              \
    do {                                                                          \
        if (threadIdx.x == 0) {                                                   \
            const long long now_ = clock64();                                     \
            atomicAdd(&g_prof[i], (unsigned long long)(now_ - prof_t_));          \
            prof_t_ = now_;                                                       \
        }                                                                         \
    } while (0)
#define PROF_INIT() long long prof_t_ = clock64()
#else
#define PROF_MARK(i) do { } while (0)
#define PROF_INIT() do { } while (0)
#endif

constexpr int D = 40;
constexpr int P = 44;               // pitch (doubles): DMMA fragment loads conflict-free
constexpr int MAT = D * P;
constexpr int ROWB = D * 8;
constexpr int K = 2 * D + 1;        // sigma points

constexpr int NTH = 128;
constexpr int NB = 5;               // 8 x 8 tile grid
// bytes the bulk copies bring in: lower block triangle of S, all of A, m and b
constexpr unsigned LOAD_BYTES = 8 * (64 + 128 + 192 + 256 + 320) + D * ROWB + 2 * ROWB;

struct EnSmem {
    double Cb[MAT];   // S (lower block triangle) -> Lt, unit lower factor of S = Lt diag(dd) Lt^T; upper part zero
    double Wb[MAT];   // Vt = Lt^-1 (unit lower); upper part zero
    double Ab[MAT];   // A(t) -> A Lt
    double mv[D], isg[D];
    double bv[D];     // b(t); after the residual phase: q (first-order weights)
    double cv[D];     // A m - b + theta; after the residual phase: d (second-order weights)
    double dd[D];     // pivots d_j
    double rp[D];     // 1 / d_j
    double sdv[D];    // sqrt(c d_j): column scale of the sigma points
    double var[K + 3];
    double esde;
    uint64_t bar;
    int bad;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// reciprocal and reciprocal square root from the hardware seed (MUFU.RCP64H / RSQ64H, ~20 bits)
// plus two Newton steps: full double precision without the long IEEE division / sqrt sequences
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

// ---- (a) diagonal 8 x 8 block kb: C_kk = Lt_kk D_k Lt_kk^T -> Lt_kk (into Cb), D_k, 1/D_k and
//      Tt_kk = Lt_kk^-1 (into Wb) ---------------------------------------------------------------
// This is the serial spine of the whole kernel (one item is latency bound by five of these), so:
// every lane of the warp holds the WHOLE lower triangle (36 values) in registers and runs the
// elimination redundantly -- no shuffles, no shared-memory round trips; per pivot the dependent
// chain is reciprocal -> multiply -> one FMA; no square roots (the LDL^T form is kept by every
// consumer); no branches: all lanes store the same values to the same addresses.  Lane c also
// solves column c of the inverse of the unit-lower factor, off the critical path.
__device__ __forceinline__ void factor_diag(EnSmem& sm, int kb, int lane)
{
    double c[8][8];
    double* tile = &sm.Cb[(8 * kb) * P + 8 * kb];
    double* tinv = &sm.Wb[(8 * kb) * P + 8 * kb];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; j += 2) {   // warp-uniform (broadcast) 16-byte loads
            const double2 v = *reinterpret_cast<const double2*>(&tile[i * P + j]);
            c[i][j] = v.x;
            if (j + 1 <= i) c[i][j + 1] = v.y;
        }
    __syncwarp();   // every lane holds the tile before it is overwritten
    const int cc = lane & 7;
    double y[8];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double pj = c[j][j];
        bad |= !(pj > 0.0);
        const double rpj = fast_rcp(pj);
        // row j+1 first: it carries the next pivot
        if (j + 1 < 8) {
            const double l1 = c[j + 1][j] * rpj;
            c[j + 1][j + 1] = fma(-l1, c[j + 1][j], c[j + 1][j + 1]);
#pragma unroll
            for (int i = 7; i > j + 1; --i) {   // descending: c[i][j] is dead for the rows that follow
                const double li = c[i][j] * rpj;
#pragma unroll
                for (int m = j + 1; m <= i; ++m) c[i][m] = fma(-li, c[m][j], c[i][m]);
                c[i][j] = li;
            }
            c[j + 1][j] = l1;
        }
        sm.dd[8 * kb + j] = pj;
        sm.rp[8 * kb + j] = rpj;
        // row j of the unit-lower factor is final: entry j of column cc of its inverse ...
        double acc = (j == cc) ? 1.0 : 0.0;
#pragma unroll
        for (int m = 0; m < j; ++m) acc = fma(-c[j][m], y[m], acc);
        y[j] = acc;
        tinv[j * P + cc] = acc;
        // ... and row j of Lt with explicit ones / zeros
#pragma unroll
        for (int m = 0; m < 8; m += 2) {
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int mm = m + e;
                v[e] = (mm < j) ? c[j][mm < j ? mm : 0] : (mm == j ? 1.0 : 0.0);
            }
            *reinterpret_cast<double2*>(&tile[j * P + m]) = make_double2(v[0], v[1]);
        }
    }
    if (bad) sm.bad = 1;
}

// one 8x8x8 tile product accumulate: acc += sum_{kk<8} Aop[m][kk] * Bop[kk][n]
// AEXPR gives A[m=g][kk] for this lane's kk = h*4+q ; BEXPR gives B[kk][n=g]
#define TILE_MMA(acc0, acc1, AEXPR, BEXPR)                    \
    do {                                                      \
        _Pragma("unroll") for (int h = 0; h < 2; ++h) {       \
            const int kk = 4 * h + q;                         \
            const double a_ = (AEXPR);                        \
            const double b_ = (BEXPR);                        \
            dmma(acc0, acc1, a_, b_);                         \
        }                                                     \
    } while (0)

// ---- (b) panel tile (i, KB): Lt_ik = C_ik Tt_kk^T D_k^-1, in place ------------------------
template <int KB>
__device__ __forceinline__ void panel_tile(EnSmem& sm, int i, int g, int q)
{
    double c0 = 0.0, c1 = 0.0;
    TILE_MMA(c0, c1, sm.Cb[(8 * i + g) * P + 8 * KB + kk], sm.Wb[(8 * KB + g) * P + 8 * KB + kk]);
    const double2 r = *reinterpret_cast<const double2*>(&sm.rp[8 * KB + 2 * q]);
    // mma.sync consumed every lane's operands: the tile may be overwritten
    *reinterpret_cast<double2*>(&sm.Cb[(8 * i + g) * P + 8 * KB + 2 * q]) = make_double2(c0 * r.x, c1 * r.y);
}

// ---- (c) trailing tile (i, j) -= Lt_ik D_k Lt_jk^T ------------------------------------------
template <int KB>
__device__ __forceinline__ void trail_tile(EnSmem& sm, int i, int j, int g, int q)
{
    double2 cc = *reinterpret_cast<double2*>(&sm.Cb[(8 * i + g) * P + 8 * j + 2 * q]);
    TILE_MMA(cc.x, cc.y, -sm.Cb[(8 * i + g) * P + 8 * KB + kk] * sm.dd[8 * KB + kk],
             sm.Cb[(8 * j + g) * P + 8 * KB + kk]);
    *reinterpret_cast<double2*>(&sm.Cb[(8 * i + g) * P + 8 * j + 2 * q]) = cc;
}

// ---- tile (I, J), I > J, of Vt = Lt^-1: Vt_IJ = -Tt_II sum_{m=J}^{I-1} Lt_Im Vt_mJ; block rows
//      < I of Vt and block columns < I of Lt are final --------------------------------------
template <int I, int J>
__device__ __forceinline__ void v_tile(EnSmem& sm, int g, int q)
{
    double s0 = 0.0, s1 = 0.0, u0 = 0.0, u1 = 0.0;
#pragma unroll
    for (int m = J; m < I; ++m) {   // two accumulator pairs: half the dependent DMMA chain
        if ((m - J) & 1) TILE_MMA(u0, u1, sm.Cb[(8 * I + g) * P + 8 * m + kk], sm.Wb[(8 * m + kk) * P + 8 * J + g]);
        else             TILE_MMA(s0, s1, sm.Cb[(8 * I + g) * P + 8 * m + kk], sm.Wb[(8 * m + kk) * P + 8 * J + g]);
    }
    *reinterpret_cast<double2*>(&sm.Wb[(8 * I + g) * P + 8 * J + 2 * q]) = make_double2(s0 + u0, s1 + u1);
    __syncwarp();
    double v0 = 0.0, v1 = 0.0;
    TILE_MMA(v0, v1, -sm.Wb[(8 * I + g) * P + 8 * I + kk], sm.Wb[(8 * I + kk) * P + 8 * J + g]);
    __syncwarp();
    *reinterpret_cast<double2*>(&sm.Wb[(8 * I + g) * P + 8 * J + 2 * q]) = make_double2(v0, v1);
}

// ---- tile (i, J) of A Lt, in place over A: column J of Lt is final, columns > J of A intact ----
template <int J>
__device__ __forceinline__ void al_tile(EnSmem& sm, int i, int g, int q)
{
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
    for (int Kb = J; Kb < NB; ++Kb) {
        dmma(c0, c1, sm.Ab[(8 * i + g) * P + 8 * Kb + q], sm.Cb[(8 * Kb + q) * P + 8 * J + g]);
        dmma(e0, e1, sm.Ab[(8 * i + g) * P + 8 * Kb + 4 + q], sm.Cb[(8 * Kb + 4 + q) * P + 8 * J + g]);
    }
    *reinterpret_cast<double2*>(&sm.Ab[(8 * i + g) * P + 8 * J + 2 * q]) = make_double2(c0 + e0, c1 + e1);
}

// row I of Vt (tiles J < I), task-distributed
template <int I, int J0 = 0>
struct VRow {
    template <typename F>
    static __device__ __forceinline__ void run(EnSmem& sm, int g, int q, F&& mine)
    {
        if constexpr (J0 < I) {
            if (mine()) v_tile<I, J0>(sm, g, q);
            VRow<I, J0 + 1>::run(sm, g, q, mine);
        }
    }
};

// one block step of the factorisation; on entry Lt_kk / Tt_kk / D_k of block KB are in place.
// While warp 0 runs the serial spine (next diagonal tile + its factorisation), warps 1-3 finish
// the trailing update and then work in its shadow: block row KB of Vt and block column KB of A Lt
// (both only need what is final by now).
template <int KB>
__device__ __forceinline__ void factor_step(EnSmem& sm, int warp, int lane, int g, int q)
{
    // (b) panel rows i = KB+1..4: one tile per warp
    if (warp < NB - 1 - KB) panel_tile<KB>(sm, KB + 1 + warp, g, q);
    __syncthreads();
    if (warp == 0) {
        trail_tile<KB>(sm, KB + 1, KB + 1, g, q);
        __syncwarp();
        factor_diag(sm, KB + 1, lane);
    } else {
        int n = 0;
        auto mine = [&]() { return (n++ % 3) == warp - 1; };
#pragma unroll
        for (int i = KB + 2; i < NB; ++i)
#pragma unroll
            for (int j = KB + 1; j <= i; ++j)
                if (mine()) trail_tile<KB>(sm, i, j, g, q);
        if constexpr (KB >= 1) VRow<KB>::run(sm, g, q, mine);
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (mine()) al_tile<KB>(sm, i, g, q);
    }
    __syncthreads();
}

// ---- tile-row I of dEsde/dS = (c/2) Vt^T diag(dw) Vt (lower tiles J <= I), mirrored on store ----
template <int I>
__device__ __forceinline__ void deds_row(const EnSmem& sm, double* __restrict__ oEs, double sc, int g, int q)
{
    double a[NB - I][2];
#pragma unroll
    for (int Kb = I; Kb < NB; ++Kb)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int kr = 8 * Kb + 4 * h + q;
            a[Kb - I][h] = sm.cv[kr] * sm.Wb[kr * P + 8 * I + g];
        }
    const int r = 8 * I + g;
#pragma unroll
    for (int J = 0; J <= I; ++J) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int Kb = I; Kb < NB; ++Kb)
#pragma unroll
            for (int h = 0; h < 2; ++h) dmma(c0, c1, a[Kb - I][h], sm.Wb[(8 * Kb + 4 * h + q) * P + 8 * J + g]);
        const int cc = 8 * J + 2 * q;
        const double v0 = sc * c0, v1 = sc * c1;
        if (I != J) {
            *reinterpret_cast<double2*>(&oEs[r * D + cc]) = make_double2(v0, v1);
            oEs[cc * D + r] = v0;
            oEs[(cc + 1) * D + r] = v1;
        } else {   // diagonal tile: keep the lower triangle, mirror it
            if (r >= cc) { oEs[r * D + cc] = v0; oEs[cc * D + r] = v0; }
            if (r >= cc + 1) { oEs[r * D + cc + 1] = v1; oEs[(cc + 1) * D + r] = v1; }
        }
    }
}

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

    PROF_INIT();
    if (tid == 0) {
        sm.bad = 0;
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    // S(t) (lower block triangle: row r brings columns 0 .. 8 (r/8 + 1) - 1), A(t), m(t), b(t)
    if (tid == 0) mbar_arrive_expect_tx(&sm.bar, LOAD_BYTES);
    if (tid < D) bulk_g2s(sm.Cb + tid * P, St + tid * D, 64 * ((tid >> 3) + 1), &sm.bar);
    else if (tid < 2 * D) bulk_g2s(sm.Ab + (tid - D) * P, At + (tid - D) * D, ROWB, &sm.bar);
    else if (tid == 2 * D) bulk_g2s(sm.mv, mt, ROWB, &sm.bar);
    else if (tid == 2 * D + 1) bulk_g2s(sm.bv, bt, ROWB, &sm.bar);
    // while the copies fly: zero the strict upper tiles of the L and V buffers (the copies do
    // not touch them), 1 / sigma
    {
        const double2 z = make_double2(0.0, 0.0);
#pragma unroll
        for (int I = 0; I < NB - 1; ++I) {
            const int npair = 4 * (NB - 1 - I);               // 16-byte pairs per row right of tile I
            for (int e = tid; e < 8 * npair; e += NTH) {
                const int r = e / npair, cp = e - r * npair;
                const int o = (8 * I + r) * P + 8 * (I + 1) + 2 * cp;
                *reinterpret_cast<double2*>(&sm.Cb[o]) = z;
                *reinterpret_cast<double2*>(&sm.Wb[o]) = z;
            }
        }
    }
    if (tid < D) sm.isg[tid] = 1.0 / b.sigma[p * b.sigma_stride + tid];
    mbar_wait(&sm.bar, 0u);
    PROF_MARK(0);

    // <f>, <df/dx> for vgpa_eval_full (lorenz_96.py:34-83,440-462); S is still intact (lower part)
    if (ex.Efx != nullptr && lp == 0) {
        for (int i = tid; i < D; i += NTH) {
            const int f1 = (i + 1) % D, b1 = (i + D - 1) % D, b2 = (i + D - 2) % D;
            const double s1 = sm.Cb[(f1 > b1 ? f1 : b1) * P + (f1 > b1 ? b1 : f1)];
            const double s2 = sm.Cb[(b2 > b1 ? b2 : b1) * P + (b2 > b1 ? b1 : b2)];
            ex.Efx[(long long)t * D + i] = (s1 - s2) + (sm.mv[f1] - sm.mv[b2]) * sm.mv[b1] - sm.mv[i] + theta;
            double* row = ex.Edf + (long long)t * D * D + (long long)i * D;
            for (int j = 0; j < D; ++j) row[j] = 0.0;
            row[i] = -1.0;
            row[f1] = sm.mv[b1];
            row[b2] = -sm.mv[b1];
            row[b1] = sm.mv[f1] - sm.mv[b2];
        }
        __syncthreads();
    }
    // The factorisation is S = Lt diag(dd) Lt^T with Lt unit lower (numpy.linalg.cholesky reads the
    // lower triangle; so do we).  chol(c S) = Lt diag(sqrt(c dd)): the sigma points are
    // m +- sdv_j Lt[:, j] with sdv_j = sqrt(c dd_j), and V = chol(S)^-1 = diag(dd^-1/2) Vt, so the
    // square roots only ever appear as per-column scalars of the consumers.
    // ---- blocked factorisation; Vt = Lt^-1 and A Lt grow in its shadow ---------------------
    if (warp == 0) factor_diag(sm, 0, lane);
    else if (tid - 32 < D) {
        // cv = A m - b + theta (A is still intact), in the shadow of the first diagonal block
        const int r = tid - 32;
        const double* ar = sm.Ab + r * P;
        double y0 = 0.0, y1 = 0.0;
#pragma unroll
        for (int k = 0; k < D; k += 2) {
            const double2 a2 = *reinterpret_cast<const double2*>(&ar[k]);
            y0 = fma(a2.x, sm.mv[k], y0);
            y1 = fma(a2.y, sm.mv[k + 1], y1);
        }
        sm.cv[r] = ((y0 + y1) - sm.bv[r]) + theta;
    }
    PROF_MARK(1);
    __syncthreads();
    factor_step<0>(sm, warp, lane, g, q);
    PROF_MARK(2);
    factor_step<1>(sm, warp, lane, g, q);
    PROF_MARK(3);
    factor_step<2>(sm, warp, lane, g, q);
    PROF_MARK(4);
    factor_step<3>(sm, warp, lane, g, q);
    PROF_MARK(5);
    // ---- what needed the last diagonal block: row 4 of Vt, column 4 of A Lt, the column scales ----
    {
        if (warp == 0) { v_tile<4, 0>(sm, g, q); al_tile<4>(sm, 4, g, q); }
        else if (warp == 1) { v_tile<4, 1>(sm, g, q); al_tile<4>(sm, 0, g, q); }
        else if (warp == 2) { v_tile<4, 2>(sm, g, q); al_tile<4>(sm, 1, g, q); al_tile<4>(sm, 2, g, q); }
        else { v_tile<4, 3>(sm, g, q); al_tile<4>(sm, 3, g, q); }
        if (tid < D) {
            const double cd = c * sm.dd[tid];
            sm.sdv[tid] = cd * fast_rsqrt(cd);
        }
    }
    PROF_MARK(6);
    __syncthreads();
    PROF_MARK(8);

    // ---- residual energies of the 81 sigma points: ONE THREAD PER SIGMA POINT walks the 40
    //      state entries with a sliding window (x[i-2], x[i-1], x[i], x[i+1]); lanes of a
    //      warp read consecutive columns of Lt and A Lt (conflict-free), the per-entry
    //      constants are warp-uniform broadcasts, and no cross-lane reduction is needed.
    //      The upper triangle of the Lt buffer is true zeros, so no selects are needed ----
    if (tid < K) {
        const int k = tid;
        const int kp = (k == 0) ? K - 1 : k - 1, kn = (k == K - 1) ? 0 : k + 1;
        const int col = (k == 0) ? 0 : ((k <= D) ? k - 1 : k - 1 - D);
        const int colp = (kp == 0) ? 0 : ((kp <= D) ? kp - 1 : kp - 1 - D);
        const int coln = (kn == 0) ? 0 : ((kn <= D) ? kn - 1 : kn - 1 - D);
        const double sg = (k == 0) ? 0.0 : ((k <= D) ? sm.sdv[col] : -sm.sdv[col]);     // +- sqrt(c d_col)
        const double sgp = (kp == 0) ? 0.0 : ((kp <= D) ? sm.sdv[colp] : -sm.sdv[colp]);
        const double sgn = (kn == 0) ? 0.0 : ((kn <= D) ? sm.sdv[coln] : -sm.sdv[coln]);
        const double* Lc = sm.Cb + col;
        const double* ALc = sm.Ab + col;
        // the flattened roll (lorenz_96.py:27-32) wraps into the neighbouring sigma points
        double xm2 = fma(sgp, sm.Cb[(D - 2) * P + colp], sm.mv[D - 2]);
        double xm1 = fma(sgp, sm.Cb[(D - 1) * P + colp], sm.mv[D - 1]);
        double x0 = fma(sg, Lc[0], sm.mv[0]);
        const double xwrap = fma(sgn, sm.Cb[coln], sm.mv[0]);
        double var = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double xp1 = (i + 1 < D) ? fma(sg, Lc[(i + 1) * P], sm.mv[i + 1]) : xwrap;
            const double fx = fma(xp1 - xm2, xm1, -x0);                 // lorenz_96.py:85-101 (theta is in cv)
            const double r = fx + fma(sg, ALc[i * P], sm.cv[i]);
            var = fma(sm.isg[i] * r, r, var);
            xm2 = xm1;
            xm1 = x0;
            x0 = xp1;
        }
        sm.var[k] = var;
    }
    __syncthreads();
    PROF_MARK(9);
    {
        // Esde(t) = 1/2 sum_k w_k var_k  (fixed order: lanes stride the 81 values; every warp
        // computes it, so no second barrier is needed before the weights)
        double e = 0.0;
        for (int k = lane; k < K; k += 32) e += (k == 0 ? w0 : wi) * sm.var[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        e *= 0.5;
        if (tid == 0) sm.esde = e;
        if (tid < D) {
            const int j = tid;
            const double vp = sm.var[1 + j], vm = sm.var[1 + D + j];
            const double dj = sm.dd[j];
            // V = diag(dd^-1/2) Vt: fold the scales into the weights
            sm.bv[j] = (wi * (vp - vm)) * fast_rsqrt(dj);                       // q_j / sqrt(d_j)
            sm.cv[j] = (0.5 * (wi * (vp + vm)) - e * (1.0 / c)) * sm.rp[j];     // d_j-weight / d_j
        }
    }
    __syncthreads();
    PROF_MARK(10);

    double* oEm = s.dEm + ((long long)lp * N + t) * D;
    double* oEs = s.dEs + ((long long)lp * N + t) * D * D;
    // ---- dEsde/dS = (c/2) V^T diag(d) V = (c/2) Vt^T diag(d / dd) Vt, lower tiles, mirrored ------
    // tile rows by cost (I + 1)(5 - I): warp0: I=2, warp1: I=3, warp2: I=1, warp3: I=0 and 4
    {
        const double sc = 0.5 * c;
        if (warp == 0) deds_row<2>(sm, oEs, sc, g, q);
        else if (warp == 1) deds_row<3>(sm, oEs, sc, g, q);
        else if (warp == 2) deds_row<1>(sm, oEs, sc, g, q);
        else {
            deds_row<0>(sm, oEs, sc, g, q);
            deds_row<4>(sm, oEs, sc, g, q);
        }
    }
    // ---- dEsde/dm = (sqrt(c)/2) V^T q = (sqrt(c)/2) Vt^T (q / sqrt(dd)); Vt is zero above the
    //      diagonal, so every thread of a warp runs the same k range ----
    if (tid < D) {
        double a0 = 0.0, a1 = 0.0;
        const int k0 = (warp == 0) ? 0 : 32;
#pragma unroll 4
        for (int k = k0; k < D; k += 2) {
            a0 = fma(sm.Wb[k * P + tid], sm.bv[k], a0);
            a1 = fma(sm.Wb[(k + 1) * P + tid], sm.bv[k + 1], a1);
        }
        oEm[tid] = (0.5 * sqrt(c)) * (a0 + a1);
    }
    if (tid == 0) {
        s.esde_t[(long long)lp * N + t] = sm.esde;
        if (sm.bad) atomicCAS(&s.status[lp], 0, 1 + t);
    }
    PROF_MARK(11);
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
