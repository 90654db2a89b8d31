// common.cuh -- shared declarations of the sm_100a VGPA kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vgpa {

enum Model { MODEL_DW = 0, MODEL_OU = 1, MODEL_L63 = 2, MODEL_L96 = 3 };
enum Method { ODE_EULER = 0, ODE_HEUN = 1, ODE_RK2 = 2, ODE_RK4 = 3 };

// Device-resident description of a problem batch (all pointers are device
// pointers; *_stride in elements, 0 = shared by every problem).
struct Batch {
    int model, method, D, N, M, B;
    double dt, dt_model;
    const double* theta;  long long theta_stride;
    const double* sigma;  long long sigma_stride;
    const double* R;      long long R_stride;
    const long long* obs_t;   // M
    const int* obs_index;     // N : ordinal of the observation at grid index t, or -1
    const double* obs_y;  long long obs_y_stride;
    const double* m0;     long long m0_stride;
    const double* s0;     long long s0_stride;
    const double* E0;     long long E0_stride;
    const int* active;        // B flags or null: problems whose flag is 0 are skipped by every kernel
                              // (their F, gradient and scratch rows are left untouched)
    const int* plist;         // or null.  Compacted launches of the D = 40 kernels (vgpa_set_active_list): the
                              // pass covers list positions [p0, p0 + count) and the problem at position k is
                              // plist[k], so that a thinned-out ensemble still fills whole waves
};

// problem served by launch position `pos` (= p0 + local index)
__device__ __forceinline__ int problem_at(const Batch& b, int pos) { return b.plist != nullptr ? b.plist[pos] : pos; }

// Per-pass (chunk) scratch: trajectories of the marginal moments, the SDE-energy
// gradients and the per-time-step energy.  Problem-major: [p][t][...].
struct Scratch {
    double* mt;      // (C, N, D)
    double* st;      // (C, N, D, D)
    double* dEm;     // (C, N, D)
    double* dEs;     // (C, N, D, D)
    double* esde_t;  // (C, N)
    int* status;     // (C) : 0 ok, else 1 + time index of the first non-PD S(t)
};

// Optional per-problem trajectory outputs (vgpa_eval_full); null = not stored.
struct Extra {
    double* lamt;  // (N, D)
    double* psit;  // (N, D, D)
    double* Efx;   // (N, D)
    double* Edf;   // (N, D, D)
    double* parts; // (3) E0, Esde, Eobs
};

// Batched sample paths / observations (datagen.cu); device pointers, strides in elements (0 = shared).
struct TrajArgs {
    int model, D, N, B;
    double dt;
    const double* theta;  long long theta_stride;    // DW [theta], OU [theta, mu], L63 [sigma, rho, beta], L96 [F]
    const double* sigma;  long long sigma_stride;    // (D) diagonal of the system noise
    const double* x_init; long long x_init_stride;   // (D) state at t0; null (L63, L96) = the reference's burn-in
    const double* z;      long long z_stride;        // (D, N) standard-normal draws, the reference's layout
    double* path;         long long path_stride;     // (N, D)
};
struct ObsArgs {
    int D, N, M, B;
    const long long* obs_t;                          // (M)
    const double* R;      long long R_stride;        // (D) diagonal of the observation noise
    const double* path;   long long path_stride;     // (N, D)
    const double* xi;     long long xi_stride;       // (D, M) standard-normal draws
    double* obs_y;        long long obs_y_stride;    // (M, D)
};

// ---- shared-memory layout of a 40 x 40 matrix in the D = 40 kernels -----------------------
// Rows are contiguous (one bulk / cp.async copy per row); row PAIRS are 84 doubles apart:
//   address(row, col) = 84 * (row >> 1) + 40 * (row & 1) + col      (1680 doubles per matrix)
// i.e. pitch 40 plus a cumulative shift of 4 doubles per row pair.  Modulo the 16 8-byte banks a
// half-warp sees, row r starts at 8 (r & 1) + 4 ((r >> 1) & 3), which makes ALL access shapes of
// the kernels bank-conflict free (64-bit accesses go out per half-warp, 128-bit per quarter-warp):
//   DMMA A fragments  (lane (g,q): row R+g,    col C+q)       rows g = 0..3 start at 0, 8, 4, 12
//   DMMA B fragments  (lane (g,q): row R+q,    col C+g)       same with q and g swapped
//   accumulator pairs (lane (g,q): row R+g,    cols C+2q, +1 as one 16-byte access)
//   transposed reads  (lane (g,q): row R+2q+e, col C+g)       rows 2q+e start at 8e + 4q
// (pitch 44 / 42, used before, made the 16-byte accumulator accesses 2-way conflicted: 22 % of the
// forward sweep's shared-memory wavefronts, profiles/README.md.)
constexpr int SM_MAT = 84 * 20;          // doubles per matrix
constexpr int SM_R8 = 84 * 4;            // offset of 8 rows
__host__ __device__ constexpr int sm_idx(int row, int col) { return 84 * (row >> 1) + 40 * (row & 1) + col; }
// B-fragment addressing: element (8 B + 4 h + q, C + g) is at  B * SM_R8 + C + sm_boff(q, g) + h * SM_BH
__host__ __device__ constexpr int sm_boff(int q, int g) { return 84 * (q >> 1) + 40 * (q & 1) + g; }
constexpr int SM_BH = 84 * 2;

// ---- launchers (one translation unit each) --------------------------------
// p0: first problem of the chunk; count: problems in the chunk.
void launch_small_fwd(const Batch& b, const Scratch& s, const double* x, long long x_stride,
                      int p0, int count, cudaStream_t st);
// D = 1 batches, F and gradient wanted, nothing else kept: the whole evaluation in one launch (small_dim.cu,
// scan1_eval_kernel).  Returns false when it does not apply; the caller then runs the four phases.
bool launch_small_fused(const Batch& b, const double* x, long long x_stride, double* F, double* grad,
                        long long grad_stride, int p0, int count, const Extra& ex, cudaStream_t st);
void launch_small_energy(const Batch& b, const Scratch& s, const double* x, long long x_stride,
                         int p0, int count, const Extra& ex, cudaStream_t st);
void launch_small_bwd(const Batch& b, const Scratch& s, const double* x, long long x_stride,
                      double* grad, long long grad_stride, int p0, int count, const Extra& ex,
                      cudaStream_t st);

// Lorenz-63 batches of up to a few thousand problems: several lanes per problem (l63_lanes.cu)
bool l63_lanes_applies(const Batch& b, int count);
void launch_l63_fwd_lanes(const Batch& b, const Scratch& s, const double* x, long long x_stride, int p0, int count,
                          cudaStream_t st);
void launch_l63_bwd_lanes(const Batch& b, const Scratch& s, const double* x, long long x_stride, double* grad,
                          long long grad_stride, int p0, int count, cudaStream_t st);

void launch_l96_fwd(const Batch& b, const Scratch& s, const double* x, long long x_stride,
                    int p0, int count, cudaStream_t st);
void launch_l96_energy(const Batch& b, const Scratch& s, const double* x, long long x_stride,
                       int p0, int count, const Extra& ex, cudaStream_t st);
void launch_l96_bwd(const Batch& b, const Scratch& s, const double* x, long long x_stride,
                    double* grad, long long grad_stride, int p0, int count, const Extra& ex,
                    cudaStream_t st);

// F[p] = E0 + trapz(esde_t) (/sigma for DW, OU) + Eobs.
void launch_finalize(const Batch& b, const Scratch& s, double* F, int p0, int count,
                     const Extra& ex, cudaStream_t st);

// Stand-alone backward sweep with dense jump tables (BwdOde.__call__).
void launch_bwd_dense(int method, int D, int N, double dt, const double* A, const double* dEm,
                      const double* dEs, const double* jm, const double* js, double* lam,
                      double* psi, cudaStream_t st);

// VarGP.initialization for problems p0 .. p0 + count - 1 (init.cu); scratch: count * D * 2 (M + 2) doubles.
void launch_initialization(const Batch& b, int p0, int count, double t0, double* scratch, double* x,
                           long long x_stride, int* err, cudaStream_t st);

// Hyper-parameter gradients of model.energy (hyper.cu); all pointers are device pointers.
void launch_hyper(int model, int D, int N, double dt_model, const double* theta, const double* sigma,
                  const double* x, const double* mt, const double* st, const double* esde, double* ft,
                  double* fs, double* dth, double* dsig, int* status, cudaStream_t stream);
// dEobs_dr of the 1-D likelihood (N values, zero except at obs_t).
void launch_obs_dr(int N, int M, const long long* obs_t, const double* obs_y, const double* R,
                   const double* mt, const double* st, double* dr, cudaStream_t stream);

// Euler-Maruyama sample paths and noisy observations for ensembles (make_trajectory / collect_obs).
void launch_trajectories(const TrajArgs& a, cudaStream_t st);
void launch_collect_obs(const ObsArgs& a, cudaStream_t st);

// Dense jump tables dEobs_dm (N,D), dEobs_ds (N,D,D) (GaussianLikelihood.gradients).
void launch_jump_tables(int D, int N, int M, const long long* obs_t, const double* obs_y, const double* R,
                        const double* mt, double* jm, double* js, cudaStream_t st);

}  // namespace vgpa
