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
//   followers (warps 1-3, one step behind, STATIC work map): panels Lt_ik = C_ik Tt_kk^T D_k^-1 as
//          DMMA tile products, trailing C_ij -= Lt_ik D_k Lt_jk^T, then in the shadow of the spine
//          block row k of Vt (a warp per block column) and block column k of A Lt for the warp's
//          own tile rows, whose A fragments stay in REGISTERS for the whole item (A m comes with
//          their load; the product lands in place over A)
//   81 residual energies: a thread serves both signs of a column, rows in three segments (123 threads);
//   Vt^T diag(w / d) Vt on the lower tiles, mirrored on store; dE/dm from the same columns.
// Shared-memory layout: common.cuh sm_idx (conflict-free for every access shape used here).
// What bounds it (profiles/README.md, round 2): no single pipe -- issue slots ~56 %, L1 data pipe
// ~65 %, FP64 pipe (DMMA + scalar, one pipe) ~53 % -- with dependent chains everywhere; the CODE SIZE
// matters as much as the instruction count (the five CTAs of an SM sit in different phases: at 47 KB
// of SASS the instruction cache hit rate fell to 87 % and two leaner variants ran 10 % slower).
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
    double varp[3][K + 3];   // residual energies of the sigma points: partial sums over the three row segments
    int bad;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// named barriers (ids 1..): producer side arrives without blocking, consumer side waits -- the PTX
// producer / consumer pattern (st.shared; bar.arrive | bar.sync; ld.shared), in which the barrier
// itself orders the producer's earlier shared-memory stores before the consumer's loads.  (A
// __threadfence_block() in front of the arrive compiles to MEMBAR.SC.CTA, ~150 cycles with stores in
// flight, twice per block on the serial spine.)  Ids and counts are immediates so that the compiler
// reserves exactly the barriers used.
template <int ID>
__device__ __forceinline__ void bar_sync()
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(128) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_arrive()
{
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(128) : "memory");
}

// reciprocal from the hardware seed (MUFU.RCP64H, ~20 bits) and ONE cubically convergent step
// r (1 + e + e^2), e = 1 - x r: three dependent FMAs instead of the four of two Newton steps (this
// sits on the pivot-to-pivot chain of the factorisation); error ~ e^3 = 2^-60 before rounding.
// Reciprocal square root: the hardware seed plus two Newton steps.
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    return fma(fma(e, e, e), r, r);
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


// ---- block column J of A Lt for the tile rows of this warp, A fragments from REGISTERS (a0: tile
//      row i0, a1: tile row i0 + 1 when `pair`), in place over A in shared memory: column J of Lt is
//      final.  The tile rows share every B fragment; the two k-halves accumulate separately
//      (half the dependent DMMA chain). ------------------------------------------------------
__device__ __forceinline__ void al_col(EnSmem& sm, const double (&a0)[NB][2], const double (&a1)[NB][2], bool pair, int i0,
                                       int J, int g, int q)
{
    double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, e[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const double* lb = &sm.Cb[8 * J + sm_boff(q, g)];             // B fragments of block column J
#pragma unroll
    for (int Kb = 0; Kb < NB; ++Kb)
        if (Kb >= J) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double bf = lb[Kb * SM_R8 + h * SM_BH];
                dmma(c[h][0], c[h][1], a0[Kb][h], bf);
                if (pair) dmma(e[h][0], e[h][1], a1[Kb][h], bf);
            }
        }
    *reinterpret_cast<double2*>(&sm.Ab[sm_idx(8 * i0 + g, 8 * J + 2 * q)]) = make_double2(c[0][0] + c[1][0], c[0][1] + c[1][1]);
    if (pair)
        *reinterpret_cast<double2*>(&sm.Ab[sm_idx(8 * i0 + 8 + g, 8 * J + 2 * q)]) =
            make_double2(e[0][0] + e[1][0], e[0][1] + e[1][1]);
}

// ---- a follower warp (1..3) for the whole factorisation, one step behind the spine.  Per step kb:
//        wait B1 (Tt_kk, Lt_kk, dd, rp of block kb are in place), my panel tile (kb + 1 + warp, kb),
//        B2 (all panels of the column are in place), my trailing tiles of the step (static map
//        `trail`: 8 bits per tile, i | j << 4, 4 tiles per step at most 3 used), B3 (arrive), and in
//        the shadow of the spine what has just become final: tile (kb, warp - 1) of Vt and block
//        column kb of A Lt for MY tile rows (A fragments in REGISTERS for the whole item: loaded
//        once, the product lands in place over A; cv = A m - b + theta comes with the load).
//      After the last diagonal block (CTA barrier): column 4 of A Lt and my tile of row 4 of Vt.
__device__ __forceinline__ void follower(EnSmem& sm, int warp, double theta, int g, int q)
{
    // static work map (one code path for the three warps: the kernel's code size is what the instruction
    // cache sees, profiles/README.md):           trailing tiles of steps 0 | 1 | 2        Vt column  A Lt rows
    //   warp 1:  (2,1) (3,2) (4,3) | (3,2) (4,4) | (4,3)                                  0          4
    //   warp 2:  (2,2) (3,3) (4,1) | (3,3) (4,2) | (4,4)                                  1          0, 1
    //   warp 3:  (3,1) (4,2) (4,4) | (4,3)       | -                                      2          2, 3
    const bool pair = warp != 1;
    const int i0 = (warp == 1) ? 4 : (warp == 2 ? 0 : 2);
    const unsigned trail0 = (warp == 1) ? 0x342312u : (warp == 2 ? 0x143322u : 0x442413u);
    const unsigned trail1 = (warp == 1) ? 0x4423u : (warp == 2 ? 0x2433u : 0x34u);
    const unsigned trail2 = (warp == 1) ? 0x34u : (warp == 2 ? 0x44u : 0u);
    double a0[NB][2], a1[NB][2];
    {
        const double* aa = &sm.Ab[sm_idx(8 * i0 + g, q)];         // row 8 i0 + 8 + g has the same skew
        const int o1 = pair ? SM_R8 : 0;
        double y0 = 0.0, y1 = 0.0;
#pragma unroll
        for (int Kb = 0; Kb < NB; ++Kb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double mk = sm.mv[8 * Kb + 4 * h + q];
                a0[Kb][h] = aa[8 * Kb + 4 * h];
                a1[Kb][h] = aa[o1 + 8 * Kb + 4 * h];
                y0 = fma(a0[Kb][h], mk, y0);
                y1 = fma(a1[Kb][h], mk, y1);
            }
        y0 += __shfl_xor_sync(0xffffffffu, y0, 1);
        y0 += __shfl_xor_sync(0xffffffffu, y0, 2);
        y1 += __shfl_xor_sync(0xffffffffu, y1, 1);
        y1 += __shfl_xor_sync(0xffffffffu, y1, 2);
        if (q == 0) {
            sm.cv[8 * i0 + g] = (y0 - sm.bv[8 * i0 + g]) + theta;
            if (pair) sm.cv[8 * i0 + 8 + g] = (y1 - sm.bv[8 * i0 + 8 + g]) + theta;
        }
    }
#pragma unroll 1
    for (int kb = 0; kb < NB; ++kb) {
        if (kb < NB - 1) {
            if (kb & 1) bar_sync<2>();                         // B1
            else bar_sync<1>();
            if (warp < NB - 1 - kb) panel_tile(sm, kb + 1 + warp, kb, g, q);
            bar_sync<3>();                                     // B2
#pragma unroll 1
            for (unsigned w = (kb == 0) ? trail0 : (kb == 1 ? trail1 : (kb == 2 ? trail2 : 0u)); w != 0u; w >>= 8)
                trail_tile(sm, (int)(w & 15u), (int)((w >> 4) & 15u), kb, g, q);
            if (kb < NB - 2) bar_arrive<4>();                  // B3
            if (warp - 1 < kb) v_tile(sm, kb, warp - 1, g, q);
        } else {
            __syncthreads();                                   // the last diagonal block is in place
        }
        al_col(sm, a0, a1, pair, i0, kb, g, q);
    }
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
    const int lp = blockIdx.x / N, t = blockIdx.x - lp * N, p = problem_at(b, p0 + lp);
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
    // thread (r0, ch) = (tid / 20, tid % 20) copies chunk ch of rows r0, r0 + 6, ...  (Bulk / TMA copies were
    // measured slower twice: 82 row copies in round 1, and 40 row-PAIR copies of 640 bytes completing on one
    // mbarrier in round 2 -- no LSU wavefronts for the load, yet 15.43 -> 15.86 ms: every bulk copy with per-lane
    // operands costs a ~15-instruction serialisation loop, and the TMA unit queues them.)
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
    // (An L2 prefetch of the item that will follow this one in its SM slot was here in round 1; without it
    // the kernel is 4 % faster: 365 instructions per item for lines that arrive in time anyway.)
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
    // (Rotating the roles with the CTA, so that the spines of the co-resident CTAs do not all sit on
    // sub-partition 0, was measured: 3 % SLOWER -- the spines run best away from the tile work.)
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
        PROF_MARK(1);
        __syncthreads();
    } else {
        follower(sm, warp, theta, g, q);
    }
    PROF_MARK(2);
    // ---- what needed the last diagonal block: row 4 of Vt (column 4 of A Lt: follower()), the column scales ----
    {
        v_tile(sm, 4, warp, g, q);
        if (tid < D) {
            const double cd = c * sm.dd[tid];
            sm.sdv[tid] = cd * fast_rsqrt(cd);
        }
    }
    PROF_MARK(6);
    __syncthreads();
    PROF_MARK(8);

    // ---- residual energies of the 81 sigma points.  One thread serves BOTH sigma points m +- s L[:, col]
    //      of a column (they share every load: the column entries of Lt and A Lt, m, cv, 1 / sigma); the
    //      centre point is a 41st "column" with s = 0.  The 40 state entries are cut into three segments
    //      (rows 0..13, 14..26, 27..39), so 41 x 3 = 123 threads each walk at most 14 rows with a sliding
    //      window (x[i-2], x[i-1], x[i], x[i+1]) per sign; the three partial sums of a sigma point are
    //      added in a fixed order below.  (One thread per sigma point over all 40 rows: 81 threads, 40
    //      dependent steps, 975 shared-memory wavefronts per item; this: 14 steps, ~600 wavefronts.)
    //      The upper triangle of the Lt buffer is true zeros: no selects. ----
    if (tid < 3 * (D + 1)) {
        const int seg = tid / (D + 1), task = tid - seg * (D + 1);
        const int i0 = (seg == 0) ? 0 : (seg == 1 ? 14 : 27), i1 = (seg == 0) ? 14 : (seg == 1 ? 27 : D);
        const bool centre = task == D;
        const int col = centre ? 0 : task;
        const double sd = centre ? 0.0 : sm.sdv[col];                           // sqrt(c d_col)
        const double* Lc = sm.Cb + col;
        const double* ALc = sm.Ab + col;
        // windows of the two signs at the first row of the segment
        double pm2, pm1, p0, qm2, qm1, q0;      // p: m + sd L[:, col] (or the centre point), q: m - sd L[:, col]
        {
            // the flattened roll (lorenz_96.py:27-32) wraps rows -2, -1 of a sigma point into the PREVIOUS
            // one: +col follows +(col-1) (+0 follows the centre), -col follows -(col-1) (-0 follows +39),
            // the centre follows -39.  Inside a segment the "previous rows" are the thread's own column.
            const int colp = (seg != 0) ? col : ((centre || col == 0) ? D - 1 : col - 1);
            const double sp = (seg != 0) ? sd : sm.sdv[colp];
            const double sgp_p = (seg != 0) ? sd : (centre ? -sp : (col == 0 ? 0.0 : sp));
            const double sgp_q = (seg != 0) ? -sd : (col == 0 ? sp : -sp);
            const int r2 = (seg != 0) ? i0 - 2 : D - 2, r1 = (seg != 0) ? i0 - 1 : D - 1;
            const double l2 = sm.Cb[sm_idx(r2, colp)], l1 = sm.Cb[sm_idx(r1, colp)], l0 = Lc[sm_idx(i0, 0)];
            pm2 = fma(sgp_p, l2, sm.mv[r2]);
            pm1 = fma(sgp_p, l1, sm.mv[r1]);
            qm2 = fma(sgp_q, l2, sm.mv[r2]);
            qm1 = fma(sgp_q, l1, sm.mv[r1]);
            p0 = fma(sd, l0, sm.mv[i0]);
            q0 = fma(-sd, l0, sm.mv[i0]);
        }
        // row 40 wraps into row 0 of the NEXT sigma point: +col -> +(col+1) (+39 -> -0), -col -> -(col+1)
        // (-39 -> the centre), the centre -> +0
        double pwrap = 0.0, qwrap = 0.0;
        if (seg == 2) {
            const int coln = (centre || col == D - 1) ? 0 : col + 1;
            const double sn = sm.sdv[coln], l0n = sm.Cb[coln];
            const double sgn_p = centre ? sn : (col == D - 1 ? -sn : sn);
            const double sgn_q = (col == D - 1) ? 0.0 : -sn;
            pwrap = fma(sgn_p, l0n, sm.mv[0]);
            qwrap = fma(sgn_q, l0n, sm.mv[0]);
        }
        double vp = 0.0, vq = 0.0;
        // one row: x[i+1] of both signs given, the residuals of row i accumulate, the windows slide
        auto row = [&](int i, double pp1, double qp1) {
            const double al = sd * ALc[sm_idx(i, 0)], cvi = sm.cv[i], isg = sm.isg[i];
            const double rp_ = fma(pp1 - pm2, pm1, -p0) + (cvi + al);           // lorenz_96.py:85-101 (theta is in cv)
            const double rq_ = fma(qp1 - qm2, qm1, -q0) + (cvi - al);
            vp = fma(isg * rp_, rp_, vp);
            vq = fma(isg * rq_, rq_, vq);
            pm2 = pm1; pm1 = p0; p0 = pp1;
            qm2 = qm1; qm1 = q0; q0 = qp1;
        };
        const int ilast = (seg == 2) ? D - 1 : i1;      // rows below ilast have their x[i+1] in the thread's own column
#pragma unroll 2
        for (int i = i0; i < ilast; ++i) {
            const double l1 = Lc[sm_idx(i + 1, 0)], m1 = sm.mv[i + 1];
            row(i, fma(sd, l1, m1), fma(-sd, l1, m1));
        }
        if (seg == 2) row(D - 1, pwrap, qwrap);
        if (centre) {
            sm.varp[seg][0] = vp;
        } else {
            sm.varp[seg][1 + col] = vp;
            sm.varp[seg][1 + D + col] = vq;
        }
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
        for (int k = lane; k < K; k += 32)
            e += (k == 0 ? w0 : wi) * ((sm.varp[0][k] + sm.varp[1][k]) + sm.varp[2][k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        e *= 0.5;
        esde = e;
        for (int j = lane; j < D; j += 32) {
            const double vp = (sm.varp[0][1 + j] + sm.varp[1][1 + j]) + sm.varp[2][1 + j];
            const double vm = (sm.varp[0][1 + D + j] + sm.varp[1][1 + D + j]) + sm.varp[2][1 + D + j];
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
