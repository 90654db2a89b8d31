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
// Blocked algorithm on 8 x 8 tiles (5 x 5 tile grid), LDL^T form: S = Lt D Lt^T with Lt unit
// lower, Vt = Lt^-1; chol(c S) = Lt diag(sqrt(c d)) and V = diag(d^-1/2) Vt, so square roots only
// appear as per-column scalars of the consumers.  Scalar FP64 instructions and DMMA share ONE pipe
// on sm_100a (a warp-wide DFMA costs 2.26 pipe cycles whatever the lane mask, a DMMA 16:
// profiles/microbench_r01.jsonl), so everything that can be a tile product is one.  One item is
// LATENCY bound by its serial spine, the five diagonal blocks (tools/energy_prof.cu), hence:
//   load : lower block triangle of S(t), A(t), m(t), b(t) by 16-byte cp.async copies; the upper
//          tiles of the Lt buffer are zero-filled while the copies are in flight
//   spine (warp 0, never waits at a CTA barrier): for k = 0..4 update the diagonal tile, factor it
//          with the whole lower triangle in the registers of every lane (redundant, branch-free;
//          lane c also solves column c of Tt_kk = Lt_kk^-1), announce it (named barrier, arrive),
//          compute its own panel tile (k+1, k)
//   followers (warps 1-3, one step behind): panels Lt_ik = C_ik Tt_kk^T D_k^-1 as DMMA tile
//          products, trailing C_ij -= Lt_ik D_k Lt_jk^T, then in the shadow of the spine block
//          row k of Vt and block column k of A Lt (in place over A; column 0 also gives A m)
//   81 residual energies, one thread per sigma point;
//   Vt^T diag(w / d) Vt on the lower tiles, mirrored on store; dE/dm from the same columns.
// Shared-memory layout: common.cuh sm_idx (conflict-free for every access shape used here).
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
constexpr int MAT = SM_MAT;         // skewed layout of common.cuh (sm_idx): every access shape conflict-free
constexpr int ROWB = D * 8;
constexpr int K = 2 * D + 1;        // sigma points

constexpr int NTH = 128;
constexpr int NB = 5;               // 8 x 8 tile grid

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
    int bad;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// named barriers (ids 1..): producer side arrives without blocking, consumer side waits
// (ids and counts are immediates so that the compiler reserves exactly the barriers used)
template <int ID>
__device__ __forceinline__ void bar_sync()
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(128) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_arrive()
{
    // producer side of the PTX producer/consumer pattern (st.shared; bar.arrive | bar.sync; ld.shared);
    // the block-scope fence makes the ordering of the preceding shared-memory stores explicit
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(128) : "memory");
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
    double* tile = &sm.Cb[kb * SM_R8 + 8 * kb];      // element (8 kb + i, 8 kb + j) = tile[sm_idx(i, j)]
    double* tinv = &sm.Wb[kb * SM_R8 + 8 * kb];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; j += 2) {   // warp-uniform (broadcast) 16-byte loads
            const double2 v = *reinterpret_cast<const double2*>(&tile[sm_idx(i, j)]);
            c[i][j] = v.x;
            if (j + 1 <= i) c[i][j + 1] = v.y;
        }
    __syncwarp();   // every lane holds the tile before it is overwritten
    const int cc = lane & 7;
    double y[8];
    bool bad = false;
    // What this lane will store at the end: chunk (lane & 3) of row (lane >> 2) of Lt, and
    // d / 1/d of pivot (lane & 7).  Captured with selects as the rows become final, so the spine
    // has no divergent code and the whole tile leaves in ONE store instruction (a warp-wide store of
    // identical values costs 4 shared-memory wavefronts, a divergent single-lane store costs the
    // spine a reconvergence: both were measured).
    const int lrow = lane >> 2, lch = lane & 3;
    double2 mine = make_double2(0.0, 0.0);
    double dmine = 0.0, rmine = 0.0;
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
        if (cc == j) {
            dmine = pj;
            rmine = rpj;
        }
        // row j of the unit-lower factor is final: entry j of column cc of its inverse ...
        double acc = (j == cc) ? 1.0 : 0.0;
#pragma unroll
        for (int m = 0; m < j; ++m) acc = fma(-c[j][m], y[m], acc);
        y[j] = acc;
        tinv[sm_idx(j, cc)] = acc;
        // ... and my chunk of row j of Lt (explicit ones / zeros)
        {
            double2 pick = make_double2(0.0, 0.0);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                double v[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int mm = 2 * ch + e;
                    v[e] = (mm < j) ? c[j][mm < j ? mm : 0] : (mm == j ? 1.0 : 0.0);
                }
                if (lch == ch) pick = make_double2(v[0], v[1]);
            }
            if (lrow == j) mine = pick;
        }
    }
    *reinterpret_cast<double2*>(&tile[sm_idx(lrow, 2 * lch)]) = mine;
    sm.dd[8 * kb + cc] = dmine;
    sm.rp[8 * kb + cc] = rmine;
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

// All tile indices below are RUN-TIME values and the block loop of the factorisation is not
// unrolled: the kernel is executed once per CTA as straight-line code, so its size is what the
// instruction cache sees (a fully unrolled version, 64 KB, hit the cache only 73 % of the time
// and starved the serial spine: profiles/README.md).
// ---- (b) panel tile (i, kb): Lt_ik = C_ik Tt_kk^T D_k^-1, in place ------------------------
__device__ __forceinline__ void panel_tile(EnSmem& sm, int i, int kb, int g, int q)
{
    double c0 = 0.0, c1 = 0.0;
    const double* ca = &sm.Cb[sm_idx(8 * i + g, 8 * kb)];
    const double* tb = &sm.Wb[sm_idx(8 * kb + g, 8 * kb)];
    TILE_MMA(c0, c1, ca[kk], tb[kk]);
    const double2 r = *reinterpret_cast<const double2*>(&sm.rp[8 * kb + 2 * q]);
    // mma.sync consumed every lane's operands: the tile may be overwritten
    *reinterpret_cast<double2*>(&sm.Cb[sm_idx(8 * i + g, 8 * kb + 2 * q)]) = make_double2(c0 * r.x, c1 * r.y);
}

// ---- (c) trailing tile (i, j) -= Lt_ik D_k Lt_jk^T ------------------------------------------
__device__ __forceinline__ void trail_tile(EnSmem& sm, int i, int j, int kb, int g, int q)
{
    double* ct = &sm.Cb[sm_idx(8 * i + g, 8 * j + 2 * q)];
    const double* la = &sm.Cb[sm_idx(8 * i + g, 8 * kb)];
    const double* lb = &sm.Cb[sm_idx(8 * j + g, 8 * kb)];
    const double* dk = &sm.dd[8 * kb];
    double2 cc = *reinterpret_cast<double2*>(ct);
    TILE_MMA(cc.x, cc.y, -la[kk] * dk[kk], lb[kk]);
    *reinterpret_cast<double2*>(ct) = cc;
}

// ---- tile (I, J), I > J, of Vt = Lt^-1: Vt_IJ = -Tt_II sum_{m=J}^{I-1} Lt_Im Vt_mJ; block rows
//      < I of Vt and block columns < I of Lt are final --------------------------------------
__device__ __forceinline__ void v_tile(EnSmem& sm, int I, int J, int g, int q)
{
    double s0 = 0.0, s1 = 0.0;
    const double* la = &sm.Cb[sm_idx(8 * I + g, 0)];
    const double* vb = &sm.Wb[8 * J + sm_boff(q, g)];           // B fragments of block column J
#pragma unroll 1
    for (int m = J; m < I; ++m) TILE_MMA(s0, s1, la[8 * m + kk], vb[m * SM_R8 + h * SM_BH]);
    double* out = &sm.Wb[sm_idx(8 * I + g, 8 * J + 2 * q)];
    *reinterpret_cast<double2*>(out) = make_double2(s0, s1);
    __syncwarp();
    double v0 = 0.0, v1 = 0.0;
    const double* ta = &sm.Wb[sm_idx(8 * I + g, 8 * I)];
    TILE_MMA(v0, v1, -ta[kk], vb[I * SM_R8 + h * SM_BH]);
    __syncwarp();
    *reinterpret_cast<double2*>(out) = make_double2(v0, v1);
}

// ---- tiles (i0, J) and (i0 + 1, J) of A Lt (PAIR: i0 + 1 < 5), in place over A: column J of Lt is
//      final, columns > J of A intact.  The two tile rows share every B fragment.  Column 0 reads
//      all of A, so cv = A m - b + theta comes with it (WITH_CV). ----------------------------
template <bool PAIR, bool WITH_CV>
__device__ __forceinline__ void al_tiles(EnSmem& sm, int i0, int J, double theta, int g, int q)
{
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, y0 = 0.0, y1 = 0.0;
    const double* aa = &sm.Ab[sm_idx(8 * i0 + g, q)];             // row 8 i0 + 8 + g has the same skew
    const double* lb = &sm.Cb[8 * J + sm_boff(q, g)];             // B fragments of block column J
#pragma unroll 1
    for (int Kb = J; Kb < NB; ++Kb) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ko = 8 * Kb + 4 * h;
            const double bf = lb[Kb * SM_R8 + h * SM_BH];
            const double a0 = aa[ko];
            dmma(c0, c1, a0, bf);
            double a1 = 0.0;
            if (PAIR) {
                a1 = aa[SM_R8 + ko];
                dmma(e0, e1, a1, bf);
            }
            if (WITH_CV) {
                const double mk = sm.mv[ko + q];
                y0 = fma(a0, mk, y0);
                if (PAIR) y1 = fma(a1, mk, y1);
            }
        }
    }
    *reinterpret_cast<double2*>(&sm.Ab[sm_idx(8 * i0 + g, 8 * J + 2 * q)]) = make_double2(c0, c1);
    if (PAIR) *reinterpret_cast<double2*>(&sm.Ab[sm_idx(8 * i0 + 8 + g, 8 * J + 2 * q)]) = make_double2(e0, e1);
    if (WITH_CV) {
        y0 += __shfl_xor_sync(0xffffffffu, y0, 1);
        y0 += __shfl_xor_sync(0xffffffffu, y0, 2);
        if (PAIR) {
            y1 += __shfl_xor_sync(0xffffffffu, y1, 1);
            y1 += __shfl_xor_sync(0xffffffffu, y1, 2);
        }
        if (q == 0) {
            sm.cv[8 * i0 + g] = (y0 - sm.bv[8 * i0 + g]) + theta;
            if (PAIR) sm.cv[8 * i0 + 8 + g] = (y1 - sm.bv[8 * i0 + 8 + g]) + theta;
        }
    }
}
// the three A Lt tasks of block column J: tile rows (0,1), (2,3), (4)
template <bool WITH_CV>
__device__ __forceinline__ void al_task(EnSmem& sm, int task, int J, double theta, int g, int q)
{
    if (task < 2) al_tiles<true, WITH_CV>(sm, 2 * task, J, theta, g, q);
    else al_tiles<false, WITH_CV>(sm, 4, J, theta, g, q);
}

// ---- tile-row I of dEsde/dS = (c/2) Vt^T diag(dw) Vt (lower tiles J <= I), mirrored on store,
//      and entries 8I..8I+7 of dEsde/dm = (sqrt(c)/2) Vt^T qw ----------------------------------
__device__ __forceinline__ void deds_row(const EnSmem& sm, double* __restrict__ oEs, double* __restrict__ oEm,
                                         double sc, double scm, int I, int g, int q)
{
    const int r = 8 * I + g;
    {   // lane (g, q): column r, rows 8I + q, +4, ... (Vt is zero above the diagonal)
        double am = 0.0;
#pragma unroll 1
        for (int kr = 8 * I + q; kr < D; kr += 4) am = fma(sm.Wb[sm_idx(kr, r)], sm.bv[kr], am);
        am += __shfl_xor_sync(0xffffffffu, am, 1);
        am += __shfl_xor_sync(0xffffffffu, am, 2);
        if (q == 0) oEm[r] = scm * am;
    }
    // A fragments dw[k] Vt[k][r], k = 8 (I + n) + q (+4): loaded once, shared by every J
    double af[NB][2];
    const int nk = NB - I;
#pragma unroll
    for (int n = 0; n < NB; ++n)
        if (n < nk) {
            const int k0 = 8 * (I + n) + q;
            af[n][0] = sm.cv[k0] * sm.Wb[(I + n) * SM_R8 + 8 * I + sm_boff(q, g)];
            af[n][1] = sm.cv[k0 + 4] * sm.Wb[(I + n) * SM_R8 + 8 * I + sm_boff(q, g) + SM_BH];
        }
#pragma unroll 1
    for (int J = 0; J <= I; ++J) {
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;   // two accumulator pairs: half the dependent chain
        const double* vb = &sm.Wb[I * SM_R8 + 8 * J + sm_boff(q, g)];
#pragma unroll
        for (int n = 0; n < NB; ++n)
            if (n < nk) {
                dmma(c0, c1, af[n][0], vb[n * SM_R8]);
                dmma(e0, e1, af[n][1], vb[n * SM_R8 + SM_BH]);
            }
        const int cc = 8 * J + 2 * q;
        const double v0 = sc * (c0 + e0), v1 = sc * (c1 + e1);
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
    if (b.active != nullptr && b.active[p] == 0) return;   // the whole CTA: before any barrier
    const double* At = x + (long long)p * xs + (long long)t * D * D;
    const double* bt = x + (long long)p * xs + (long long)N * D * D + (long long)t * D;
    const double* mt = s.mt + ((long long)lp * N + t) * D;
    const double* St = s.st + ((long long)lp * N + t) * D * D;
    const double theta = b.theta[p * b.theta_stride];
    const double kap = 1.05 * D, c = D + kap;                 // utilities.py:271
    const double w0 = kap / c, wi = 1.0 / (2.0 * c);          // :290-291

    PROF_INIT();
    if (tid == 0) sm.bad = 0;
    // S(t) (lower block triangle), A(t), m(t), b(t) by 16-byte cp.async copies (SASS LDGSTS):
    // thread (r0, ch) = (tid / 20, tid % 20) copies chunk ch of rows r0, r0 + 6, ... (82 small
    // bulk copies per CTA were measured slower: the TMA unit serialises them)
    if (tid < 120) {
        const int r0 = tid / 20, ch = tid - r0 * 20;
        const double* sg_ = St + r0 * D + 2 * ch;
        const double* ag_ = At + r0 * D + 2 * ch;
#pragma unroll
        for (int n = 0; n < 7; ++n) {
            const int row = r0 + 6 * n;
            if (row < D) {
                const int o = sm_idx(row, 2 * ch);
                if (ch < 4 * ((row >> 3) + 1)) cp_async16(sm.Cb + o, sg_ + 6 * n * D);
                cp_async16(sm.Ab + o, ag_ + 6 * n * D);
            }
        }
    } else {
        const int u = tid - 120;   // 8 threads: the two 40-vectors (20 chunks each)
#pragma unroll
        for (int n = 0; n < 5; ++n) {
            const int c2 = u + 8 * n;
            if (c2 < 20) cp_async16(sm.mv + 2 * c2, mt + 2 * c2);
            else cp_async16(sm.bv + 2 * (c2 - 20), bt + 2 * (c2 - 20));
        }
    }
    cp_async_commit();
    // pull the inputs of the item that will follow this one in its SM slot (148 SMs x 5 CTAs
    // further down the grid) into L2: its load phase then sees L2 instead of HBM latency
    {
        constexpr int AHEAD = 148 * 5;
        const long long nb = (long long)blockIdx.x + AHEAD;
        if (nb < (long long)gridDim.x) {
            if (tid >= 2 * D && tid < 3 * D) {
                const int r = tid - 2 * D;
                bulk_prefetch_l2(s.st + nb * (D * D) + r * D, 64 * ((r >> 3) + 1));
            } else if (tid == 3 * D) {
                const int lpn = (int)(nb / N), tn = (int)(nb - (long long)lpn * N);
                bulk_prefetch_l2(x + (long long)(p0 + lpn) * xs + (long long)tn * D * D, D * ROWB);
            }
        }
    }
    // while the copies fly: zero the strict upper tiles of the L buffer (the copies do not touch
    // them; the residual phase reads whole columns), 1 / sigma
    {
        const double2 z = make_double2(0.0, 0.0);
#pragma unroll
        for (int I = 0; I < NB - 1; ++I) {
            const int npair = 4 * (NB - 1 - I);               // 16-byte pairs per row right of tile I
            for (int e = tid; e < 8 * npair; e += NTH) {
                const int r = e / npair, cp = e - r * npair;
                const int o = sm_idx(8 * I + r, 8 * (I + 1) + 2 * cp);
                *reinterpret_cast<double2*>(&sm.Cb[o]) = z;
            }
        }
    }
    if (tid < D) sm.isg[tid] = 1.0 / b.sigma[p * b.sigma_stride + tid];
    cp_async_wait<0>();
    __syncthreads();
    PROF_MARK(0);

    // <f>, <df/dx> for vgpa_eval_full (lorenz_96.py:34-83,440-462); S is still intact (lower part)
    if (ex.Efx != nullptr && lp == 0) {
        for (int i = tid; i < D; i += NTH) {
            const int f1 = (i + 1) % D, b1 = (i + D - 1) % D, b2 = (i + D - 2) % D;
            const double s1 = sm.Cb[sm_idx(f1 > b1 ? f1 : b1, f1 > b1 ? b1 : f1)];
            const double s2 = sm.Cb[sm_idx(b2 > b1 ? b2 : b1, b2 > b1 ? b1 : b2)];
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
    // Block loop (not unrolled).  Warp 0 runs the serial spine and NEVER waits for the others
    // except for data it needs: per block kb it updates the diagonal tile, factors it, announces it
    // (named barrier B1, arrive), computes its own panel tile (kb+1, kb) (announced on B2) and goes
    // on.  Warps 1-3 follow one step behind: wait for B1, their panel tiles, B2 (all panels of the
    // column are in place), the trailing update of step kb (announced on B3, which warp 0 checks
    // just before it touches tiles of the next column), and then, in the shadow of the spine, what
    // has just become final: block row kb of Vt and block column kb of A Lt.
    if (warp == 0) {
#pragma unroll 1
        for (int kb = 0; kb < NB; ++kb) {
            if (kb > 0) {
                trail_tile(sm, kb, kb, kb - 1, g, q);
                __syncwarp();
            }
            factor_diag(sm, kb, lane);
            if (kb < NB - 1) {
                if (kb & 1) bar_arrive<2>();                   // B1 (ids 1, 2 alternate): Tt_kk, Lt_kk, dd, rp
                else bar_arrive<1>();                          //     of block kb are in place
                if (kb > 0) bar_sync<4>();                     // B3: trailing update of step kb-1 complete
                panel_tile(sm, kb + 1, kb, g, q);
                bar_arrive<3>();                               // B2: my panel tile is in place
                __syncwarp();
            }
        }
    } else {
#pragma unroll 1
        for (int kb = 0; kb < NB - 1; ++kb) {
            if (kb & 1) bar_sync<2>();                         // B1
            else bar_sync<1>();
            if (warp < NB - 1 - kb) panel_tile(sm, kb + 1 + warp, kb, g, q);
            bar_sync<3>();                                     // B2
            int n = 3 - warp;           // round-robin over warps 1..3: a task is mine when n hits 3
            for (int i = kb + 2; i < NB; ++i)
                for (int j = kb + 1; j <= i; ++j)
                    if (++n == 3) { n = 0; trail_tile(sm, i, j, kb, g, q); }
            if (kb < NB - 2) bar_arrive<4>();                  // B3
            for (int j = 0; j < kb; ++j)
                if (++n == 3) { n = 0; v_tile(sm, kb, j, g, q); }
            // A Lt, block column kb: three tasks (tile rows (0,1), (2,3), (4)); column 0 also gives cv
            for (int task = 0; task < 3; ++task)
                if (++n == 3) {
                    n = 0;
                    if (kb == 0) al_task<true>(sm, task, kb, theta, g, q);
                    else al_task<false>(sm, task, kb, theta, g, q);
                }
        }
    }
    PROF_MARK(1);
    __syncthreads();
    PROF_MARK(2);
    // ---- what needed the last diagonal block: row 4 of Vt, column 4 of A Lt, the column scales ----
    {
        v_tile(sm, 4, warp, g, q);
        if (warp < 3) al_task<false>(sm, warp, 4, theta, g, q);
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
        double xm2 = fma(sgp, sm.Cb[sm_idx(D - 2, colp)], sm.mv[D - 2]);
        double xm1 = fma(sgp, sm.Cb[sm_idx(D - 1, colp)], sm.mv[D - 1]);
        double x0 = fma(sg, Lc[0], sm.mv[0]);
        const double xwrap = fma(sgn, sm.Cb[coln], sm.mv[0]);
        double var = 0.0;
#pragma unroll 8
        for (int i = 0; i < D; ++i) {
            const double xp1 = (i + 1 < D) ? fma(sg, Lc[sm_idx(i + 1, 0)], sm.mv[i + 1]) : xwrap;
            const double fx = fma(xp1 - xm2, xm1, -x0);                 // lorenz_96.py:85-101 (theta is in cv)
            const double r = fx + fma(sg, ALc[sm_idx(i, 0)], sm.cv[i]);
            var = fma(sm.isg[i] * r, r, var);
            xm2 = xm1;
            xm1 = x0;
            x0 = xp1;
        }
        sm.var[k] = var;
    }
    __syncthreads();
    PROF_MARK(9);
    double* oEm = s.dEm + ((long long)lp * N + t) * D;
    double* oEs = s.dEs + ((long long)lp * N + t) * D * D;
    double esde;
    {
        // Esde(t) = 1/2 sum_k w_k var_k  (fixed order: lanes stride the 81 values).  EVERY warp
        // computes it and the 40 + 40 weights below and stores them (same values, same addresses),
        // so only a warp-level sync separates this from the tile products that read them.
        double e = 0.0;
        for (int k = lane; k < K; k += 32) e += (k == 0 ? w0 : wi) * sm.var[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        e *= 0.5;
        esde = e;
        for (int j = lane; j < D; j += 32) {
            const double vp = sm.var[1 + j], vm = sm.var[1 + D + j];
            // V = diag(dd^-1/2) Vt: fold the scales into the weights
            sm.bv[j] = (wi * (vp - vm)) * fast_rsqrt(sm.dd[j]);                 // q_j / sqrt(d_j)
            sm.cv[j] = (0.5 * (wi * (vp + vm)) - e * (1.0 / c)) * sm.rp[j];     // d_j-weight / d_j
        }
        __syncwarp();
    }
    PROF_MARK(10);
    // ---- dEsde/dS = (c/2) V^T diag(d) V = (c/2) Vt^T diag(d / dd) Vt, lower tiles, mirrored;
    //      dEsde/dm = (sqrt(c)/2) V^T q ------
    // tile rows by cost (I + 1)(5 - I): warp0: I=2, warp1: I=3, warp2: I=1, warp3: I=0 and 4
    {
        const double sc = 0.5 * c, scm = 0.5 * sqrt(c);
        const int I = (warp == 0) ? 2 : (warp == 1 ? 3 : (warp == 2 ? 1 : 0));
        deds_row(sm, oEs, oEm, sc, scm, I, g, q);
        if (warp == 3) deds_row(sm, oEs, oEm, sc, scm, 4, g, q);
    }
    if (tid == 0) {
        s.esde_t[(long long)lp * N + t] = esde;
        if (sm.bad) atomicCAS(&s.status[lp], 0, 1 + t);
    }
    PROF_MARK(11);
}

}  // namespace

void launch_l96_energy(const Batch& b, const Scratch& s, const double* x, long long xs, int p0, int count,
                       const Extra& ex, cudaStream_t st)
{
    size_t sh = sizeof(EnSmem);
#ifdef VGPA_EN_PROF
    if (const char* e = getenv("VGPA_EN_EXTRA_SMEM")) sh += (size_t)atoi(e);   // occupancy experiments
#endif
    cudaFuncSetAttribute(l96_energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    const unsigned grid = (unsigned)((long long)count * b.N);
    l96_energy_kernel<<<grid, NTH, sh, st>>>(b, s, x, xs, p0, count, ex);
}

}  // namespace vgpa
