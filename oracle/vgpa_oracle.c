/*
 * vgpa_oracle.c -- CPU restatement (plain C, FP64) of the VGPA free-energy +
 * gradient evaluation.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It exists so that the CUDA
 * path can be checked on a box where the (Python) reference is not present.
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference
 * legs of bench.py may load it.  The product path (vgpa_b200/) never calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * below against outputs of the unmodified reference (vrettasm/VGPA) recorded
 * in tests/golden/*.npz by tests/golden/make_golden.py.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the reference root).  The restatement is literal where the arithmetic
 * matters, including the reference's quirks (the DW
 * 8*E6 coefficient, the scalar z0'z0 broadcast of the n-D prior).  Scope
 * limits shared with the CUDA path: diagonal system noise Sigma, diagonal
 * observation noise R, identity observation operator (what the sim_params
 * JSON schema can express).
 *
 * Layouts (all C-contiguous float64):
 *   x  = [A (N,D,D) | b (N,D)]     mt (N,D)   st (N,D,D)   D=1 -> same, 1x1
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

enum { MODEL_DW = 0, MODEL_OU = 1, MODEL_L63 = 2, MODEL_L96 = 3 };
enum { ODE_EULER = 0, ODE_HEUN = 1, ODE_RK2 = 2, ODE_RK4 = 3 };

typedef struct {
    int model, method;
    int D, N, M;
    double dt;            /* step of the ODE sweeps (JSON Time-window.dt)          */
    double dt_model;      /* model.time_step = |tk[1]-tk[0]| (trapz, grad scaling) */
    const double *theta;  /* DW,OU,L96: 1 value; L63: 3 values                      */
    const double *sigma;  /* D : diagonal of the system noise                        */
    const double *R;      /* D : diagonal of the observation noise                   */
    const long long *obs_t; /* M observation indices (sorted, unique)               */
    const double *obs_y;  /* M x D                                                   */
    const double *m0;     /* D                                                       */
    const double *s0;     /* D x D                                                   */
    double E0;            /* prior KL at t=0 (constant; prior_kl0.py)                */
} oracle_problem;

#define LOG2PI 1.8378770664093453 /* log(2*pi), gaussian_like.py:12 */

/* ------------------------------------------------------------------------ */
/* small dense helpers                                                       */
/* ------------------------------------------------------------------------ */
static void matmul(int D, const double *A, const double *B, double *C) /* C = A B */
{
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += A[i * D + k] * B[k * D + j];
            C[i * D + j] = s;
        }
}
static void matmul_nt(int D, const double *A, const double *B, double *C) /* C = A B^T */
{
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += A[i * D + k] * B[j * D + k];
            C[i * D + j] = s;
        }
}
static void matmul_tn(int D, const double *A, const double *B, double *C) /* C = A^T B */
{
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += A[k * D + i] * B[k * D + j];
            C[i * D + j] = s;
        }
}
static void matvec(int D, const double *A, const double *v, double *o) /* o = A v */
{
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
        for (int k = 0; k < D; ++k) s += A[i * D + k] * v[k];
        o[i] = s;
    }
}
/* lower Cholesky of the LOWER triangle of X (numpy.linalg.cholesky semantics,
 * utilities.py:104,211,275).  Returns 0, or 1 if X is not positive definite. */
static int chol_lower(int D, const double *X, double *L)
{
    memset(L, 0, sizeof(double) * D * D);
    for (int j = 0; j < D; ++j) {
        double d = X[j * D + j];
        for (int k = 0; k < j; ++k) d -= L[j * D + k] * L[j * D + k];
        if (!(d > 0.0)) return 1;
        d = sqrt(d);
        L[j * D + j] = d;
        for (int i = j + 1; i < D; ++i) {
            double s = X[i * D + j];
            for (int k = 0; k < j; ++k) s -= L[i * D + k] * L[j * D + k];
            L[i * D + j] = s / d;
        }
    }
    return 0;
}
/* solve (L L^T) u = r */
static void chol_solve(int D, const double *L, const double *r, double *u)
{
    for (int i = 0; i < D; ++i) {
        double s = r[i];
        for (int k = 0; k < i; ++k) s -= L[i * D + k] * u[k];
        u[i] = s / L[i * D + i];
    }
    for (int i = D - 1; i >= 0; --i) {
        double s = u[i];
        for (int k = i + 1; k < D; ++k) s -= L[k * D + i] * u[k];
        u[i] = s / L[i * D + i];
    }
}
/* chol_inv (utilities.py:203-237): c_inv = L^-1, x_inv = c_inv^T c_inv */
static int chol_inv(int D, const double *X, double *Xinv, double *work /* 2 D*D */)
{
    double *L = work, *Li = work + D * D;
    if (chol_lower(D, X, L)) return 1;
    memset(Li, 0, sizeof(double) * D * D);
    for (int j = 0; j < D; ++j)      /* column j of L^-1 by forward substitution */
        for (int i = j; i < D; ++i) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) s -= L[i * D + k] * Li[k * D + j];
            Li[i * D + j] = s / L[i * D + i];
        }
    matmul_tn(D, Li, Li, Xinv);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* ODE right-hand sides: src/numerics/ode_solver.py:31-95                    */
/* ------------------------------------------------------------------------ */
/* fun_mt :44   -A m + b */
static void fun_mt(int D, const double *m, const double *A, const double *b, double *o)
{
    matvec(D, A, m, o);
    for (int i = 0; i < D; ++i) o[i] = -o[i] + b[i];
}
/* fun_st :60   -A S - S A^T + Sigma  (both products formed, as the reference) */
static void fun_st(int D, const double *S, const double *A, const double *sig,
                   double *o, double *w /* D*D */)
{
    matmul(D, A, S, o);
    matmul_nt(D, S, A, w);
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j)
            o[i * D + j] = -o[i * D + j] - w[i * D + j] + (i == j ? sig[i] : 0.0);
}
/* fun_lam :77   -dE/dm + lam.dot(A^T) = -g + A lam */
static void fun_lam(int D, const double *g, const double *A, const double *lam, double *o)
{
    matvec(D, A, lam, o);
    for (int i = 0; i < D; ++i) o[i] = -g[i] + o[i];
}
/* fun_psi :94   -dE/dS + Psi A + A^T Psi */
static void fun_psi(int D, const double *G, const double *A, const double *P,
                    double *o, double *w)
{
    matmul(D, P, A, o);
    matmul_tn(D, A, P, w);
    for (int i = 0; i < D * D; ++i) o[i] = -G[i] + o[i] + w[i];
}
static void axpy_to(int n, const double *y, double a, const double *x, double *o)
{ /* o = y + a x */
    for (int i = 0; i < n; ++i) o[i] = y[i] + a * x[i];
}
static void mid_to(int n, const double *p, const double *q, double *o)
{ /* 0.5 * (p + q): runge_kutta2.py:74-75,134-136, runge_kutta4.py:72-73,146-148 */
    for (int i = 0; i < n; ++i) o[i] = 0.5 * (p[i] + q[i]);
}

/* ------------------------------------------------------------------------ */
/* forward sweep: euler.py:27-92, heun.py:28-111, runge_kutta2.py:25-102,     */
/* runge_kutta4.py:25-113 (via fwd_ode.py:45-65)                             */
/* ------------------------------------------------------------------------ */
int oracle_fwd(const oracle_problem *p, const double *x, double *mt, double *st)
{
    const int D = p->D, N = p->N, DD = D * D;
    const double dt = p->dt, h = 0.5 * dt;
    const double *A = x, *b = x + (size_t)N * DD;
    double *w = (double *)malloc(sizeof(double) * (12 * DD + 12 * D));
    double *k1 = w, *k2 = k1 + DD, *k3 = k2 + DD, *k4 = k3 + DD, *tmp = k4 + DD,
           *amid = tmp + DD, *scr = amid + DD;
    double *v1 = scr + DD, *v2 = v1 + D, *v3 = v2 + D, *v4 = v3 + D, *vt = v4 + D,
           *bmid = vt + D;
    memcpy(mt, p->m0, sizeof(double) * D);
    memcpy(st, p->s0, sizeof(double) * DD);
    for (int k = 0; k < N - 1; ++k) {
        const double *Ak = A + (size_t)k * DD, *Ap = Ak + DD;
        const double *bk = b + (size_t)k * D, *bp = bk + D;
        const double *mk = mt + (size_t)k * D, *Sk = st + (size_t)k * DD;
        double *mn = mt + (size_t)(k + 1) * D, *Sn = st + (size_t)(k + 1) * DD;
        switch (p->method) {
        case ODE_EULER: /* euler.py:84-87 */
            fun_mt(D, mk, Ak, bk, v1);
            axpy_to(D, mk, dt, v1, mn);
            fun_st(D, Sk, Ak, p->sigma, k1, scr);
            axpy_to(DD, Sk, dt, k1, Sn);
            break;
        case ODE_HEUN: /* heun.py:91-106 */
            fun_mt(D, mk, Ak, bk, v1);
            axpy_to(D, mk, dt, v1, vt);
            fun_mt(D, vt, Ap, bp, v2);
            for (int i = 0; i < D; ++i) mn[i] = mk[i] + h * (v1[i] + v2[i]);
            fun_st(D, Sk, Ak, p->sigma, k1, scr);
            axpy_to(DD, Sk, dt, k1, tmp);
            fun_st(D, tmp, Ap, p->sigma, k2, scr);
            for (int i = 0; i < DD; ++i) Sn[i] = Sk[i] + h * (k1[i] + k2[i]);
            break;
        case ODE_RK2: /* runge_kutta2.py:92,96 -- the inner covariance stage is
                         fun_st(sk, sk, sigma): S stands in for A. */
            mid_to(DD, Ak, Ap, amid);
            mid_to(D, bk, bp, bmid);
            fun_mt(D, mk, Ak, bk, v1);
            axpy_to(D, mk, h, v1, vt);
            fun_mt(D, vt, amid, bmid, v2);
            axpy_to(D, mk, dt, v2, mn);
            fun_st(D, Sk, Sk, p->sigma, k1, scr);
            axpy_to(DD, Sk, h, k1, tmp);
            fun_st(D, tmp, amid, p->sigma, k2, scr);
            axpy_to(DD, Sk, dt, k2, Sn);
            break;
        case ODE_RK4: /* runge_kutta4.py:93-108 */
            mid_to(DD, Ak, Ap, amid);
            mid_to(D, bk, bp, bmid);
            fun_mt(D, mk, Ak, bk, v1);
            axpy_to(D, mk, h, v1, vt);
            fun_mt(D, vt, amid, bmid, v2);
            axpy_to(D, mk, h, v2, vt);
            fun_mt(D, vt, amid, bmid, v3);
            axpy_to(D, mk, dt, v3, vt);
            fun_mt(D, vt, Ap, bp, v4);
            for (int i = 0; i < D; ++i)
                mn[i] = mk[i] + dt * (v1[i] + 2.0 * (v2[i] + v3[i]) + v4[i]) / 6.0;
            fun_st(D, Sk, Ak, p->sigma, k1, scr);
            axpy_to(DD, Sk, h, k1, tmp);
            fun_st(D, tmp, amid, p->sigma, k2, scr);
            axpy_to(DD, Sk, h, k2, tmp);
            fun_st(D, tmp, amid, p->sigma, k3, scr);
            axpy_to(DD, Sk, dt, k3, tmp);
            fun_st(D, tmp, Ap, p->sigma, k4, scr);
            for (int i = 0; i < DD; ++i)
                Sn[i] = Sk[i] + dt * (k1[i] + 2.0 * (k2[i] + k3[i]) + k4[i]) / 6.0;
            break;
        default:
            free(w);
            return 1;
        }
    }
    free(w);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* backward sweep: euler.py:94-154, heun.py:113-190, runge_kutta2.py:104-194, */
/* runge_kutta4.py:115-211 (via bwd_ode.py:45-65)                            */
/* lam[N-1] = psi[N-1] = 0; the jump added at each step is the one at t-1.   */
/* ------------------------------------------------------------------------ */
int oracle_bwd(const oracle_problem *p, const double *x, const double *dEm,
               const double *dEs, const double *jm, const double *js,
               double *lam, double *psi)
{
    const int D = p->D, N = p->N, DD = D * D;
    const double dt = p->dt, h = 0.5 * dt;
    const double *A = x;
    double *w = (double *)malloc(sizeof(double) * (9 * DD + 8 * D));
    double *k1 = w, *k2 = k1 + DD, *k3 = k2 + DD, *k4 = k3 + DD, *tmp = k4 + DD,
           *amid = tmp + DD, *gmid = amid + DD, *scr = gmid + DD;
    double *v1 = scr + DD, *v2 = v1 + D, *v3 = v2 + D, *v4 = v3 + D, *vt = v4 + D,
           *vmid = vt + D;
    memset(lam, 0, sizeof(double) * (size_t)N * D);
    memset(psi, 0, sizeof(double) * (size_t)N * DD);
    for (int t = N - 1; t > 0; --t) {
        const double *At = A + (size_t)t * DD, *Am = At - DD;
        const double *gt = dEm + (size_t)t * D, *gm = gt - D;
        const double *Gt = dEs + (size_t)t * DD, *Gm = Gt - DD;
        const double *lt = lam + (size_t)t * D, *Pt = psi + (size_t)t * DD;
        double *ln = lam + (size_t)(t - 1) * D, *Pn = psi + (size_t)(t - 1) * DD;
        const double *jmn = jm + (size_t)(t - 1) * D, *jsn = js + (size_t)(t - 1) * DD;
        switch (p->method) {
        case ODE_EULER: /* euler.py:146-149 */
            fun_lam(D, gt, At, lt, v1);
            for (int i = 0; i < D; ++i) ln[i] = lt[i] - v1[i] * dt + jmn[i];
            fun_psi(D, Gt, At, Pt, k1, scr);
            for (int i = 0; i < DD; ++i) Pn[i] = Pt[i] - k1[i] * dt + jsn[i];
            break;
        case ODE_HEUN: /* heun.py:170-185 */
            fun_lam(D, gt, At, lt, v1);
            axpy_to(D, lt, -dt, v1, vt);
            fun_lam(D, gm, Am, vt, v2);
            for (int i = 0; i < D; ++i) ln[i] = lt[i] - h * (v1[i] + v2[i]) + jmn[i];
            fun_psi(D, Gt, At, Pt, k1, scr);
            axpy_to(DD, Pt, -dt, k1, tmp);
            fun_psi(D, Gm, Am, tmp, k2, scr);
            for (int i = 0; i < DD; ++i) Pn[i] = Pt[i] - h * (k1[i] + k2[i]) + jsn[i];
            break;
        case ODE_RK2: /* runge_kutta2.py:180-189 */
            mid_to(DD, Am, At, amid);
            mid_to(D, gm, gt, vmid);
            mid_to(DD, Gm, Gt, gmid);
            fun_lam(D, gt, At, lt, v1);
            axpy_to(D, lt, -h, v1, vt);
            fun_lam(D, vmid, amid, vt, v2);
            for (int i = 0; i < D; ++i) ln[i] = lt[i] - dt * v2[i] + jmn[i];
            fun_psi(D, Gt, At, Pt, k1, scr);
            axpy_to(DD, Pt, -h, k1, tmp);
            fun_psi(D, gmid, amid, tmp, k2, scr);
            for (int i = 0; i < DD; ++i) Pn[i] = Pt[i] - dt * k2[i] + jsn[i];
            break;
        case ODE_RK4: /* runge_kutta4.py:191-206 */
            mid_to(DD, Am, At, amid);
            mid_to(D, gm, gt, vmid);
            mid_to(DD, Gm, Gt, gmid);
            fun_lam(D, gt, At, lt, v1);
            axpy_to(D, lt, -h, v1, vt);
            fun_lam(D, vmid, amid, vt, v2);
            axpy_to(D, lt, -h, v2, vt);
            fun_lam(D, vmid, amid, vt, v3);
            axpy_to(D, lt, -dt, v3, vt);
            fun_lam(D, gm, Am, vt, v4);
            for (int i = 0; i < D; ++i)
                ln[i] = lt[i] - dt * (v1[i] + 2.0 * (v2[i] + v3[i]) + v4[i]) / 6.0 + jmn[i];
            fun_psi(D, Gt, At, Pt, k1, scr);
            axpy_to(DD, Pt, -h, k1, tmp);
            fun_psi(D, gmid, amid, tmp, k2, scr);
            axpy_to(DD, Pt, -h, k2, tmp);
            fun_psi(D, gmid, amid, tmp, k3, scr);
            axpy_to(DD, Pt, -dt, k3, tmp);
            fun_psi(D, Gm, Am, tmp, k4, scr);
            for (int i = 0; i < DD; ++i)
                Pn[i] = Pt[i] - dt * (k1[i] + 2.0 * (k2[i] + k3[i]) + k4[i]) / 6.0 + jsn[i];
            break;
        default:
            free(w);
            return 1;
        }
    }
    free(w);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* my_trapz: utilities.py:144-201 -- scipy trapezoid summed piecewise         */
/* between the observation indices.                                          */
/* ------------------------------------------------------------------------ */
static double seg_trapz(const double *f, int lo, int hi, double dx)
{ /* trapezoid(f[lo:hi+1], dx) = sum(dx * (f[i+1] + f[i]) / 2) */
    double s = 0.0;
    for (int i = lo; i < hi; ++i) s += dx * (f[i + 1] + f[i]) / 2.0;
    return s;
}
double oracle_trapz(const double *f, int N, double dx, const long long *obs_t, int M)
{
    double tot = 0.0;
    int first = 0;
    for (int k = 0; k < M; ++k) {
        tot += seg_trapz(f, first, (int)obs_t[k], dx);
        first = (int)obs_t[k];
    }
    if (first != N - 1) tot += seg_trapz(f, first, N - 1, dx);
    return tot;
}

/* ------------------------------------------------------------------------ */
/* Double well: double_well.py:169-260 with gaussian_moments.py:43-183        */
/* ------------------------------------------------------------------------ */
static void energy_dw(const oracle_problem *p, const double *a, const double *b,
                      const double *m, const double *s, double *esde_t, double *Ef,
                      double *Edf, double *dm, double *ds)
{
    const double th = p->theta[0], sig = p->sigma[0];
    for (int t = 0; t < p->N; ++t) {
        const double mm = m[t], v = s[t], bb = b[t];
        const double c = 4.0 * th + a[t], c2 = c * c;
        const double m2 = mm * mm, m3 = m2 * mm, m4 = m2 * m2, m5 = m4 * mm, m6 = m3 * m3;
        const double v2 = v * v, v3 = v2 * v;
        /* gaussian_moments.py:57-74 */
        const double E2 = m2 + v, E3 = m3 + 3 * mm * v, E4 = m4 + 6 * m2 * v + 3 * v2;
        const double E6 = m6 + 15 * m4 * v + 45 * m2 * v2 + 15 * v3;
        /* :214 -- note the 8.0*Ex6 (the derivative below uses 16): as the reference */
        esde_t[t] = 8.0 * (E6 - c * E4 + bb * E3) + (c2 * E2) - (2.0 * bb * c * mm) + bb * bb;
        Ef[t] = 4.0 * (th * mm - E3);   /* :220 */
        Edf[t] = 4.0 * (th - 3.0 * E2); /* :223 */
        /* gaussian_moments.py:108-122 (dm) and :155-167 (ds) */
        const double Dm2 = 2 * mm, Dm3 = 3 * (m2 + v), Dm4 = 4 * (m3 + 3 * mm * v);
        const double Dm6 = 6 * (m5 + 10 * m3 * v + 15 * mm * v2);
        const double Ds2 = 1.0, Ds3 = 3 * mm, Ds4 = 6 * (m2 + v);
        const double Ds6 = 15 * m4 + 90 * m2 * v + 45 * v2;
        dm[t] = 0.5 * (16.0 * Dm6 - 8.0 * c * Dm4 + 8.0 * bb * Dm3 + c2 * Dm2 - 2.0 * bb * c) / sig; /* :243 */
        ds[t] = 0.5 * (16.0 * Ds6 - 8.0 * c * Ds4 + 8.0 * bb * Ds3 + c2 * Ds2) / sig;               /* :248 */
    }
}

/* Ornstein-Uhlenbeck: ornstein_uhlenbeck.py:165-232 */
static void energy_ou(const oracle_problem *p, const double *a, const double *b,
                      const double *m, const double *s, double *esde_t, double *Ef,
                      double *Edf, double *dm, double *ds)
{
    const double th = p->theta[0], sig = p->sigma[0];
    for (int t = 0; t < p->N; ++t) {
        const double mm = m[t], E2 = mm * mm + s[t];
        const double q1 = (th - a[t]) * (th - a[t]), q2 = a[t] * b[t];
        esde_t[t] = E2 * q1 + 2.0 * mm * (th - a[t]) * b[t] + b[t] * b[t]; /* :205 */
        Ef[t] = -th * mm;                                                   /* :211 */
        Edf[t] = -th;                                                       /* :214 */
        dm[t] = (mm * q1 + th * b[t] - q2) / sig;                           /* :217 */
        ds[t] = 0.5 * q1 / sig;                                             /* :221 */
    }
}

/* ------------------------------------------------------------------------ */
/* Lorenz 63: lorenz_63.py:237-346 (loop) and :348-568 (energy_dm_ds)          */
/* Reads the UPPER triangle of S for the energy (:388-390) and S[2,0], S[1,0] */
/* for Ef (:320-321).                                                        */
/* ------------------------------------------------------------------------ */
static void l63_step(const double *th, const double *iS, const double *at,
                     const double *bt, const double *mt, const double *st,
                     double *esde, double *Ef, double *Edf, double *dEm, double *dEs)
{
    const double vS = th[0], vR = th[1], vB = th[2];
    const double A11 = at[0], A12 = at[1], A13 = at[2], A21 = at[3], A22 = at[4],
                 A23 = at[5], A31 = at[6], A32 = at[7], A33 = at[8];
    const double b1 = bt[0], b2 = bt[1], b3 = bt[2];
    const double mx = mt[0], my = mt[1], mz = mt[2];
    const double Sxx = st[0], Sxy = st[1], Sxz = st[2], Syy = st[4], Syz = st[5], Szz = st[8];
    /* 2nd order :393-398 */
    const double Exx = Sxx + mx * mx, Exy = Sxy + mx * my, Exz = Sxz + mx * mz;
    const double Eyy = Syy + my * my, Eyz = Syz + my * mz, Ezz = Szz + mz * mz;
    /* 3rd order :401-405 */
    const double Exxy = Sxx * my + 2 * Sxy * mx + (mx * mx) * my;
    const double Exxz = Sxx * mz + 2 * Sxz * mx + (mx * mx) * mz;
    const double Exyy = Syy * mx + 2 * Sxy * my + (my * my) * mx;
    const double Exzz = Szz * mx + 2 * Sxz * mz + (mz * mz) * mx;
    const double Exyz = Sxy * mz + Sxz * my + Syz * mx + mx * my * mz;
    /* 4th order :408-411 */
    const double Exxyy = Sxx * (my * my + Syy) + Syy * (mx * mx) + 4.0 * Sxy * mx * my +
                         (mx * my) * (mx * my) + 2 * (Sxy * Sxy);
    const double Exxzz = Sxx * (mz * mz + Szz) + Szz * (mx * mx) + 4.0 * Sxz * mx * mz +
                         (mx * mz) * (mx * mz) + 2 * (Sxz * Sxz);
    /* :414-432 */
    const double EX = (vS * vS) * (Eyy + Exx - 2 * Exy) + (A11 * A11) * Exx + (A12 * A12) * Eyy +
                      (A13 * A13) * Ezz + b1 * b1 +
                      2 * (A11 * A12 * Exy + A11 * A13 * Exz - b1 * A11 * mx + A12 * A13 * Eyz -
                           b1 * A12 * my - b1 * A13 * mz +
                           vS * (A11 * Exy + A12 * Eyy + A13 * Eyz - b1 * my - A11 * Exx -
                                 A12 * Exy - A13 * Exz + b1 * mx));
    const double EY = (vR * vR) * Exx + Eyy + Exxzz + (A21 * A21) * Exx + (A22 * A22) * Eyy +
                      (A23 * A23) * Ezz + b2 * b2 +
                      2 * (Exyz - A21 * Exy - A22 * Eyy - A23 * Eyz - A21 * Exxz - A22 * Exyz -
                           A23 * Exzz + A21 * A22 * Exy + A21 * A23 * Exz + A22 * A23 * Eyz -
                           vR * (Exy + Exxz - A21 * Exx - A22 * Exy - A23 * Exz) -
                           b2 * (vR * mx - my - Exz + A21 * mx + A22 * my + A23 * mz));
    const double EZ = Exxyy + (vB * vB) * Ezz + (A31 * A31) * Exx + (A32 * A32) * Eyy +
                      (A33 * A33) * Ezz + b3 * b3 +
                      2 * (A31 * Exxy + A32 * Exyy + A33 * Exyz + A31 * A32 * Exy +
                           A31 * A33 * Exz + A32 * A33 * Eyz -
                           vB * (Exyz + A31 * Exz + A32 * Eyz + A33 * Ezz) -
                           b3 * (Exy - vB * mz + A31 * mx + A32 * my + A33 * mz));
    *esde = 0.5 * (iS[0] * EX + iS[1] * EY + iS[2] * EZ); /* :310 */

    /* derivatives of the expectations :440-487 */
    const double dExx_dmx = 2.0 * mx, dExy_dmx = my, dExz_dmx = mz;
    const double dEyy_dmy = 2.0 * my, dExy_dmy = mx, dEyz_dmy = mz;
    const double dEzz_dmz = 2.0 * mz, dExz_dmz = mx, dEyz_dmz = my;
    const double dExxy_dmx = 2.0 * Exy, dExxz_dmx = 2.0 * Exz, dExyy_dmx = Eyy,
                 dExzz_dmx = Ezz, dExyz_dmx = Eyz;
    const double dExxy_dmy = Exx, dExyy_dmy = 2.0 * Exy, dExyz_dmy = Exz;
    const double dExxz_dmz = Exx, dExzz_dmz = 2.0 * Exz, dExyz_dmz = Exy;
    const double dExxy_dSxx = my, dExxz_dSxx = mz, dExxy_dSxy = 2.0 * mx,
                 dExyy_dSxy = 2.0 * my, dExyz_dSxy = mz, dExzz_dSxz = 2.0 * mz,
                 dExyz_dSxz = my, dExxz_dSxz = 2.0 * mx, dExyy_dSyy = mx, dExyz_dSyz = mx,
                 dExzz_dSzz = mx;
    const double dExxyy_dmx = 2.0 * Exyy, dExxzz_dmx = 2.0 * Exzz, dExxyy_dmy = 2.0 * Exxy,
                 dExxzz_dmz = 2.0 * Exxz;
    const double dExxyy_dSxx = Eyy, dExxzz_dSxx = Ezz, dExxyy_dSxy = 4.0 * Exy;
    const double dExxzz_dSxz = 4.0 * Exz, dExxyy_dSyy = Exx, dExxzz_dSzz = Exx;

    /* :490-526 */
    const double dmx1 = dExx_dmx * (vS * vS + A11 * A11) +
                        2 * (dExy_dmx * (-vS * vS + vS * A11 - vS * A12 + A11 * A12) +
                             dExz_dmx * (A11 - vS) * A13 - vS * A11 * dExx_dmx + b1 * (vS - A11));
    const double dmx2 = dExxzz_dmx + dExx_dmx * (vR * vR + A21 * A21) +
                        2 * (dExy_dmx * (-vR + vR * A22 - A21 + A21 * A22) +
                             dExz_dmx * (vR * A23 + b2 + A21 * A23) + dExyz_dmx * (1 - A22) -
                             vR * dExxz_dmx + vR * A21 * dExx_dmx - A21 * dExxz_dmx -
                             A23 * dExzz_dmx - b2 * (vR + A21));
    const double dmx3 = dExxyy_dmx + (A31 * A31) * dExx_dmx +
                        2 * (dExy_dmx * (A31 * A32 - b3) + dExz_dmx * (A33 - vB) * A31 +
                             dExyz_dmx * (A33 - vB) + A31 * dExxy_dmx + A32 * dExyy_dmx - A31 * b3);
    const double dmy1 = dEyy_dmy * (vS * vS + A12 * A12) +
                        2 * (dExy_dmy * (-(vS * vS) + vS * A11 - vS * A12 + A11 * A12) +
                             dEyz_dmy * (vS + A12) * A13 + vS * A12 * dEyy_dmy - b1 * (vS + A12));
    const double dmy2 = dEyy_dmy * (1 + A22 * A22) +
                        2 * (dExy_dmy * (-vR + vR * A22 - A21 + A21 * A22) + dExyz_dmy * (1 - A22) -
                             A22 * dEyy_dmy + dEyz_dmy * (A22 * A23 - A23) + b2 * (1 - A22));
    const double dmy3 = dExxyy_dmy + (A32 * A32) * dEyy_dmy +
                        2 * (dExyz_dmy * (A33 - vB) + A31 * dExxy_dmy + A32 * dExyy_dmy +
                             dExy_dmy * (A31 * A32 - b3) + dEyz_dmy * (A33 - vB) * A32 - A32 * b3);
    const double dmz1 = (A13 * A13) * dEzz_dmz +
                        2 * (dEyz_dmz * (vS + A12) + dExz_dmz * (A11 - vS) - b1) * A13;
    const double dmz2 = dExxzz_dmz + (A23 * A23) * dEzz_dmz +
                        2 * (dExxz_dmz * (-vR - A21) + dExz_dmz * (vR * A23 + b2 + A21 * A23) +
                             dExyz_dmz * (1 - A22) + dEyz_dmz * (A22 * A23 - A23) -
                             A23 * (dExzz_dmz + b2));
    const double dmz3 = dEzz_dmz * (vB * vB + A33 * A33) +
                        2 * ((A33 - vB) * (dExyz_dmz + dExz_dmz * A31 + dEyz_dmz * A32 - b3) -
                             vB * A33 * dEzz_dmz);
    /* :529-531 */
    dEm[0] = 0.5 * (dmx1 * iS[0] + dmx2 * iS[1] + dmx3 * iS[2]);
    dEm[1] = 0.5 * (dmy1 * iS[0] + dmy2 * iS[1] + dmy3 * iS[2]);
    dEm[2] = 0.5 * (dmz1 * iS[0] + dmz2 * iS[1] + dmz3 * iS[2]);

    /* :537-561 (the d(2nd order)/dS factors are all 1) */
    const double iSx = iS[0], iSy = iS[1], iSz = iS[2];
    const double dSxx = iSx * ((vS - A11) * (vS - A11)) +
                        iSy * (dExxzz_dSxx + ((vR + A21) * (vR + A21)) - 2 * dExxz_dSxx * (vR + A21)) +
                        iSz * (dExxyy_dSxx + (A31 * A31) + 2 * A31 * dExxy_dSxx);
    const double dSxy = iSx * 2 * (vS * A11 - vS * vS - vS * A12 + A11 * A12) +
                        iSy * 2 * ((vR * A22 - vR - A21 + A21 * A22) + dExyz_dSxy * (1 - A22)) +
                        iSz * (dExxyy_dSxy + 2 * (dExyz_dSxy * (A33 - vB) + A31 * dExxy_dSxy +
                                                  A32 * dExyy_dSxy + (A31 * A32 - b3)));
    const double dSxz = iSx * 2 * (A11 - vS) * A13 +
                        iSy * (dExxzz_dSxz + 2 * ((vR * A23 + b2 + A21 * A23) + dExyz_dSxz * (1 - A22) -
                                                  dExxz_dSxz * (vR + A21) - A23 * dExzz_dSxz)) +
                        iSz * 2 * ((A33 - vB) * A31 + dExyz_dSxz * (A33 - vB));
    const double dSyy = iSx * ((vS + A12) * (vS + A12)) + iSy * ((1 - A22) * (1 - A22)) +
                        iSz * (dExxyy_dSyy + (A32 * A32) + 2 * A32 * dExyy_dSyy);
    const double dSyz = iSx * 2 * (vS + A12) * A13 +
                        iSy * 2 * (dExyz_dSyz * (1 - A22) + (A22 - 1) * A23) +
                        iSz * 2 * (dExyz_dSyz * (A33 - vB) + (A33 - vB) * A32);
    const double dSzz = iSx * (A13 * A13) + iSy * (dExxzz_dSzz + (A23 * A23) - 2 * A23 * dExzz_dSzz) +
                        iSz * ((vB - A33) * (vB - A33));
    /* :564-566 */
    dEs[0] = 0.5 * dSxx; dEs[1] = 0.5 * dSxy; dEs[2] = 0.5 * dSxz;
    dEs[3] = 0.5 * dSxy; dEs[4] = 0.5 * dSyy; dEs[5] = 0.5 * dSyz;
    dEs[6] = 0.5 * dSxz; dEs[7] = 0.5 * dSyz; dEs[8] = 0.5 * dSzz;

    /* :319-326 */
    Ef[0] = vS * (my - mx);
    Ef[1] = vR * mx - my - st[6] - mx * mz;
    Ef[2] = st[3] + mx * my - vB * mz;
    Edf[0] = -vS;     Edf[1] = vS;  Edf[2] = 0.0;
    Edf[3] = vR - mz; Edf[4] = -1.0; Edf[5] = -mx;
    Edf[6] = my;      Edf[7] = mx;  Edf[8] = -vB;
}

/* ------------------------------------------------------------------------ */
/* Lorenz 96: lorenz_96.py:316-438 with ut_approx (utilities.py:239-310),     */
/* grad_Esde_dm_ds (variational.py:339-400), l96/shift_vectors                */
/* (lorenz_96.py:27-32,85-101), E96_drift (:440-462), E96_drift_dx (:34-83).  */
/* The y_cov product of ut_approx (utilities.py:302-306) is discarded by both */
/* call sites (lorenz_96.py:398,410) and is not formed here.                  */
/* ------------------------------------------------------------------------ */
static int l96_step(int D, double theta, const double *iS /* D */, const double *at,
                    const double *bt, const double *mt, const double *st, double *esde,
                    double *Ef, double *Edf, double *dEm, double *dEs, double *w)
{
    const int K = 2 * D + 1;
    const double kap = 1.05 * D, c = D + kap;              /* utilities.py:271 */
    const double w0 = kap / c, wi = 1.0 / (2.0 * c);       /* :290-291 */
    double *cS = w, *L = cS + D * D, *chi = L + D * D, *f = chi + (size_t)K * D,
           *var = f + (size_t)K * D, *u = var + K, *v = u + D, *Sinv = v + D,
           *LS = Sinv + D * D, *wk = LS + D * D /* 2 D*D */, *acc = wk + 2 * D * D,
           *z = acc + D * D, *r = z + D;
    for (int i = 0; i < D * D; ++i) cS[i] = c * st[i];
    if (chol_lower(D, cS, L)) {
        /* utilities.py:276-279: fall back to chol(diag(S)), NOT scaled by (D+k). */
        memset(cS, 0, sizeof(double) * D * D);
        for (int i = 0; i < D; ++i) cS[i * D + i] = st[i * D + i];
        if (chol_lower(D, cS, L)) return 2;
    }
    /* sigma points :283-288: chi = [m; m + rows of L^T; m - rows of L^T] */
    for (int j = 0; j < D; ++j) chi[j] = mt[j];
    for (int k = 0; k < D; ++k)
        for (int j = 0; j < D; ++j) {
            chi[(size_t)(1 + k) * D + j] = mt[j] + L[j * D + k];
            chi[(size_t)(1 + D + k) * D + j] = mt[j] - L[j * D + k];
        }
    /* l96 on the 2-D sigma-point matrix: numba's np.roll flattens */
    const int T = K * D;
    for (int i = 0; i < T; ++i) {
        const double fw1 = chi[(i + 1) % T], bw1 = chi[(i - 1 + T) % T], bw2 = chi[(i - 2 + T) % T];
        f[i] = (fw1 - bw2) * bw1 - chi[i] + theta;
    }
    /* x_mat = (l96(chi) + chi A^T - b)^2 ; var = diag_inv_sigma . x_mat^T
     * (lorenz_96.py:380-381, variational.py:373-374) */
    double ebar = 0.0;
    for (int k = 0; k < K; ++k) {
        double vk = 0.0;
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
            for (int j = 0; j < D; ++j) s += chi[(size_t)k * D + j] * at[i * D + j];
            const double rr = f[(size_t)k * D + i] + s - bt[i];
            vk += iS[i] * (rr * rr);
        }
        var[k] = vk;
        ebar += (k == 0 ? w0 : wi) * vk;
    }
    /* Esde(t) = 0.5 * diag_inv_sig . m_bar (:401), m_bar = weights . y */
    *esde = 0.5 * ebar;

    /* gradient integrand (variational.py:376-396), UT-averaged (lorenz_96.py:410) */
    if (chol_lower(D, st, LS)) return 2;          /* np.linalg.solve(st, .) / chol_inv(st) */
    if (chol_inv(D, st, Sinv, wk)) return 2;      /* variational.py:380 */
    memset(acc, 0, sizeof(double) * D * D);
    for (int i = 0; i < D; ++i) dEm[i] = 0.0;
    for (int k = 0; k < K; ++k) {
        const double wk_ = (k == 0 ? w0 : wi);
        /* dmt_k = solve(st, var_k * chi_k) */
        for (int i = 0; i < D; ++i) r[i] = var[k] * chi[(size_t)k * D + i];
        chol_solve(D, LS, r, u);
        for (int i = 0; i < D; ++i) dEm[i] += wk_ * 0.5 * u[i];
        /* dst_k = var_k * solve(st, z z^T) . inv_st = var_k * (S^-1 z)(inv_st z)^T */
        for (int i = 0; i < D; ++i) z[i] = chi[(size_t)k * D + i] - mt[i];
        chol_solve(D, LS, z, u);
        matvec(D, Sinv, z, v);
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) acc[i * D + j] += wk_ * 0.5 * (var[k] * u[i] * v[j]);
    }
    /* :414  dEsde_dm = dmS[:D] - Esde * solve(st, mt) */
    chol_solve(D, LS, mt, u);
    for (int i = 0; i < D; ++i) dEm[i] -= (*esde) * u[i];
    /* :417-418 dEsde_dS = 0.5 * (dmS[D:] - Esde * solve(st, I)) */
    for (int i = 0; i < D * D; ++i) dEs[i] = 0.5 * (acc[i] - (*esde) * Sinv[i]);

    /* E96_drift :440-462 and E96_drift_dx :34-83 (ordinary cyclic indices) */
    memset(Edf, 0, sizeof(double) * D * D);
    for (int k = 0; k < D; ++k) {
        const int f1 = (k + 1) % D, b1 = (k - 1 + D) % D, b2 = (k - 2 + D) % D;
        Ef[k] = (st[f1 * D + b1] - st[b2 * D + b1]) + (mt[f1] - mt[b2]) * mt[b1] - mt[k] + theta;
        Edf[k * D + k] = -1.0;
        Edf[k * D + f1] = mt[b1];
        Edf[k * D + b2] = -mt[b1];
        Edf[k * D + b1] = mt[f1] - mt[b2];
    }
    return 0;
}

/* model.energy(A, b, m, S, obs_t): returns Esde (scalar) and fills Ef (N,D),
 * Edf (N,D,D), dEsde_dm (N,D), dEsde_ds (N,D,D).  esde_t (N) is scratch/out. */
int oracle_energy(const oracle_problem *p, const double *x, const double *mt,
                  const double *st, double *Esde, double *esde_t, double *Ef, double *Edf,
                  double *dEm, double *dEs)
{
    const int D = p->D, N = p->N, DD = D * D;
    const double *A = x, *b = x + (size_t)N * DD;
    int rc = 0;
    if (p->model == MODEL_DW) {
        energy_dw(p, A, b, mt, st, esde_t, Ef, Edf, dEm, dEs);
        /* double_well.py:217 */
        *Esde = 0.5 * oracle_trapz(esde_t, N, p->dt_model, p->obs_t, p->M) / p->sigma[0];
        return 0;
    }
    if (p->model == MODEL_OU) {
        energy_ou(p, A, b, mt, st, esde_t, Ef, Edf, dEm, dEs);
        /* ornstein_uhlenbeck.py:208 */
        *Esde = 0.5 * oracle_trapz(esde_t, N, p->dt_model, p->obs_t, p->M) / p->sigma[0];
        return 0;
    }
    double *iS = (double *)malloc(sizeof(double) * D);
    for (int i = 0; i < D; ++i) iS[i] = 1.0 / p->sigma[i];
    if (p->model == MODEL_L63) {
        if (D != 3) { free(iS); return 1; }
        for (int t = 0; t < N; ++t)
            l63_step(p->theta, iS, A + (size_t)t * 9, b + (size_t)t * 3, mt + (size_t)t * 3,
                     st + (size_t)t * 9, esde_t + t, Ef + (size_t)t * 3, Edf + (size_t)t * 9,
                     dEm + (size_t)t * 3, dEs + (size_t)t * 9);
    } else if (p->model == MODEL_L96) {
        const int K = 2 * D + 1;
        const size_t wsz = (size_t)8 * DD + (size_t)2 * K * D + K + 6 * D;
#pragma omp parallel
        {
            double *w = (double *)malloc(sizeof(double) * wsz);
#pragma omp for schedule(static)
            for (int t = 0; t < N; ++t) {
                int e = l96_step(D, p->theta[0], iS, A + (size_t)t * DD, b + (size_t)t * D,
                                 mt + (size_t)t * D, st + (size_t)t * DD, esde_t + t,
                                 Ef + (size_t)t * D, Edf + (size_t)t * DD, dEm + (size_t)t * D,
                                 dEs + (size_t)t * DD, w);
                if (e) {
#pragma omp atomic write
                    rc = e;
                }
            }
            free(w);
        }
    } else {
        free(iS);
        return 1;
    }
    free(iS);
    /* lorenz_63.py:336 / lorenz_96.py:428 */
    *Esde = oracle_trapz(esde_t, N, p->dt_model, p->obs_t, p->M);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* Hyper-parameter gradients of model.energy (the last two entries of its     */
/* third return value; VarGP discards them, variational.py:175):              */
/*   DW  double_well.py:251-257      OU  ornstein_uhlenbeck.py:223-229         */
/*   L63 lorenz_63.py:327-343 with Efg_drift_theta :572-633                    */
/*   L96 lorenz_96.py:420-434                                                  */
/* dth: 1 (DW, OU), 3 (L63), D (L96) values; dsig: 1 value (D = 1) or D x D.   */
/* ------------------------------------------------------------------------ */
static void l63_theta_step(const double *th, const double *at, const double *bt,
                           const double *mt, const double *st, double *V)
{ /* lorenz_63.py:587-633 (reads the UPPER triangle of S) */
    const double vS = th[0], vR = th[1], vB = th[2];
    const double A11 = at[0], A12 = at[1], A13 = at[2], A21 = at[3], A22 = at[4],
                 A23 = at[5], A31 = at[6], A32 = at[7], A33 = at[8];
    const double b1 = bt[0], b2 = bt[1], b3 = bt[2];
    const double mx = mt[0], my = mt[1], mz = mt[2];
    const double Sxx = st[0], Sxy = st[1], Sxz = st[2], Syy = st[4], Syz = st[5], Szz = st[8];
    const double Exx = Sxx + mx * mx, Exy = Sxy + mx * my, Eyy = Syy + my * my;
    const double Exz = Sxz + mx * mz, Ezz = Szz + mz * mz, Eyz = Syz + my * mz;
    const double Exxz = Sxx * mz + 2 * Sxz * mx + (mx * mx) * mz;
    const double Exyz = Sxy * mz + Sxz * my + Syz * mx + mx * my * mz;
    V[0] = Eyy * (vS + A12) + Exx * (vS - A11) + Exy * (A11 - 2 * vS - A12) +
           A13 * (Eyz - Exz) + b1 * (mx - my);
    V[1] = vR * Exx - Exy - Exxz + A21 * Exx + A22 * Exy + A23 * Exz - b2 * mx;
    V[2] = -Exyz + vB * Ezz - A31 * Exz - A32 * Eyz - A33 * Ezz + b3 * mz;
}

/* per-component residual second moments of Lorenz 63, Efg = (EX, EY, EZ) (lorenz_63.py:414-432) */
static void l63_efg_step(const double *th, const double *at, const double *bt,
                         const double *mt, const double *st, double *E)
{
    const double iS1[3] = {1.0, 0.0, 0.0}, iS2[3] = {0.0, 1.0, 0.0}, iS3[3] = {0.0, 0.0, 1.0};
    double e, Ef[3], Edf[9], dEm[3], dEs[9];
    /* esde = 0.5 * iS . Efg: unit weights pick the components */
    l63_step(th, iS1, at, bt, mt, st, &e, Ef, Edf, dEm, dEs); E[0] = 2.0 * e;
    l63_step(th, iS2, at, bt, mt, st, &e, Ef, Edf, dEm, dEs); E[1] = 2.0 * e;
    l63_step(th, iS3, at, bt, mt, st, &e, Ef, Edf, dEm, dEs); E[2] = 2.0 * e;
}

/* m_bar of Lorenz 96: UT mean of the squared residual components (lorenz_96.py:398, 423) */
static int l96_mbar_step(int D, double theta, const double *at, const double *bt, const double *mt,
                         const double *st, double *mbar, double *w)
{
    const int K = 2 * D + 1;
    const double kap = 1.05 * D, c = D + kap;
    const double w0 = kap / c, wi = 1.0 / (2.0 * c);
    double *cS = w, *L = cS + D * D, *chi = L + D * D, *f = chi + (size_t)K * D;
    for (int i = 0; i < D * D; ++i) cS[i] = c * st[i];
    if (chol_lower(D, cS, L)) {
        memset(cS, 0, sizeof(double) * D * D);
        for (int i = 0; i < D; ++i) cS[i * D + i] = st[i * D + i];
        if (chol_lower(D, cS, L)) return 2;
    }
    for (int j = 0; j < D; ++j) chi[j] = mt[j];
    for (int k = 0; k < D; ++k)
        for (int j = 0; j < D; ++j) {
            chi[(size_t)(1 + k) * D + j] = mt[j] + L[j * D + k];
            chi[(size_t)(1 + D + k) * D + j] = mt[j] - L[j * D + k];
        }
    const int T = K * D;
    for (int i = 0; i < T; ++i) {
        const double fw1 = chi[(i + 1) % T], bw1 = chi[(i - 1 + T) % T], bw2 = chi[(i - 2 + T) % T];
        f[i] = (fw1 - bw2) * bw1 - chi[i] + theta;
    }
    for (int i = 0; i < D; ++i) mbar[i] = 0.0;
    for (int k = 0; k < K; ++k)
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
            for (int j = 0; j < D; ++j) s += chi[(size_t)k * D + j] * at[i * D + j];
            const double rr = f[(size_t)k * D + i] + s - bt[i];
            mbar[i] += (k == 0 ? w0 : wi) * (rr * rr);
        }
    return 0;
}

int oracle_energy_hyper(const oracle_problem *p, const double *x, const double *mt,
                        const double *st, double *dth, double *dsig)
{
    const int D = p->D, N = p->N, DD = D * D;
    const double *A = x, *b = x + (size_t)N * DD;
    if (p->model == MODEL_DW || p->model == MODEL_OU) {
        double *e = (double *)malloc(sizeof(double) * N), *g = (double *)malloc(sizeof(double) * N);
        const double th = p->theta[0], sig = p->sigma[0];
        for (int t = 0; t < N; ++t) {
            const double mm = mt[t], v = st[t], m2 = mm * mm;
            const double E2 = m2 + v;
            if (p->model == MODEL_DW) {
                const double c = 4.0 * th + A[t], bb = b[t];
                const double E3 = m2 * mm + 3 * mm * v, E4 = m2 * m2 + 6 * m2 * v + 3 * v * v;
                const double E6 = m2 * m2 * m2 + 15 * m2 * m2 * v + 45 * m2 * v * v + 15 * v * v * v;
                e[t] = 8.0 * (E6 - c * E4 + bb * E3) + (c * c * E2) - (2.0 * bb * c * mm) + bb * bb;
                g[t] = c * E2 - 4.0 * E4 - bb * mm;                       /* :251 */
            } else {
                e[t] = E2 * (th - A[t]) * (th - A[t]) + 2.0 * mm * (th - A[t]) * b[t] + b[t] * b[t];
                g[t] = E2 * (th - A[t]) + mm * b[t];                      /* :223 */
            }
        }
        const double Esde = 0.5 * oracle_trapz(e, N, p->dt_model, p->obs_t, p->M) / sig;
        const double tz = oracle_trapz(g, N, p->dt_model, p->obs_t, p->M);
        dth[0] = (p->model == MODEL_DW ? 4.0 * tz : tz) / sig;           /* DW :251-252, OU :223-224 */
        dsig[0] = -Esde / sig;                                           /* DW :255, OU :227 */
        free(e); free(g);
        return 0;
    }
    if (p->model != MODEL_L63 && p->model != MODEL_L96) return 1;
    if (p->model == MODEL_L63 && D != 3) return 1;
    const int K = 2 * D + 1;
    double *ft = (double *)malloc(sizeof(double) * (size_t)N * D);   /* per-t theta integrand */
    double *fs = (double *)malloc(sizeof(double) * (size_t)N * D);   /* per-t sigma integrand */
    double *w = (double *)malloc(sizeof(double) * ((size_t)2 * DD + (size_t)2 * K * D));
    double *col = (double *)malloc(sizeof(double) * N);
    int rc = 0;
    for (int t = 0; t < N && !rc; ++t) {
        const double *at = A + (size_t)t * DD, *bt = b + (size_t)t * D, *m = mt + (size_t)t * D,
                     *S = st + (size_t)t * DD;
        if (p->model == MODEL_L63) {
            l63_theta_step(p->theta, at, bt, m, S, ft + (size_t)t * 3);
            l63_efg_step(p->theta, at, bt, m, S, fs + (size_t)t * 3);
        } else {
            rc = l96_mbar_step(D, p->theta[0], at, bt, m, S, fs + (size_t)t * D, w);
            for (int k = 0; k < D; ++k) {   /* Ef[t] + mt.dot(at.T) - bt  (lorenz_96.py:420) */
                const int f1 = (k + 1) % D, b1 = (k - 1 + D) % D, b2 = (k - 2 + D) % D;
                double am = 0.0;
                for (int j = 0; j < D; ++j) am += m[j] * at[k * D + j];
                const double Ef = (S[f1 * D + b1] - S[b2 * D + b1]) + (m[f1] - m[b2]) * m[b1] - m[k] + p->theta[0];
                ft[(size_t)t * D + k] = Ef + am - bt[k];
            }
        }
    }
    memset(dsig, 0, sizeof(double) * DD);
    for (int i = 0; i < D && !rc; ++i) {
        for (int t = 0; t < N; ++t) col[t] = ft[(size_t)t * D + i];
        dth[i] = (1.0 / p->sigma[i]) * oracle_trapz(col, N, p->dt_model, p->obs_t, p->M);
        for (int t = 0; t < N; ++t) col[t] = fs[(size_t)t * D + i];
        /* -0.5 inv_sigma . diag(trapz) . inv_sigma with a diagonal Sigma */
        dsig[i * D + i] = -0.5 * (1.0 / p->sigma[i]) * oracle_trapz(col, N, p->dt_model, p->obs_t, p->M) *
                          (1.0 / p->sigma[i]);
    }
    free(ft); free(fs); free(w); free(col);
    return rc;
}

/* dEobs_dr of GaussianLikelihood.gradients: 1-D (gaussian_like.py:194) at the observation
 * indices, zero elsewhere; the n-D branch (:226) never fills it (zeros (N, M, M)). */
void oracle_eobs_dr(const oracle_problem *p, const double *mt, const double *st, double *dr)
{
    if (p->D != 1) {
        memset(dr, 0, sizeof(double) * (size_t)p->N * p->M * p->M);
        return;
    }
    memset(dr, 0, sizeof(double) * p->N);
    for (int n = 0; n < p->M; ++n) {
        const long long t = p->obs_t[n];
        const double y = p->obs_y[n], m = mt[t], Ex2 = m * m + st[t];
        dr[t] = -0.5 * ((y * y) - 2.0 * y * m + Ex2 + 1.0) / p->R[0];
    }
}

/* ------------------------------------------------------------------------ */
/* VarGP.initialization (variational.py:73-139): x0 = [A0 | b0] from a cubic    */
/* spline through the observations, each dimension separately.  The spline is  */
/* scipy.interpolate.CubicSpline with its default 'not-a-knot' ends (scipy is   */
/* an unpinned dependency of the reference, requirements.txt; 1.18.1 in the     */
/* authoring container): first derivatives s at the knots from the tridiagonal  */
/* system of CubicSpline.__init__ (n >= 4), the parabola special case (n = 3),  */
/* the straight line (n = 2); then the Hermite cubic of CubicHermiteSpline,     */
/* evaluated as c3 + c2 d + c1 d^2 + c0 d^3.  tw[k] = t0 + k * dt_model.        */
/* ------------------------------------------------------------------------ */
static int spline_slopes(int n, const double *x, const double *y, double *s, double *w /* 2 n */)
{
    if (n < 2) return 1;
    for (int i = 0; i + 1 < n; ++i)
        if (!(x[i + 1] > x[i])) return 1;                       /* scipy: x must be strictly increasing */
    if (n == 2) { s[0] = s[1] = (y[1] - y[0]) / (x[1] - x[0]); return 0; }
    if (n == 3) { /* parabola through the three points: 3 x 3 system */
        const double dx0 = x[1] - x[0], dx1 = x[2] - x[1];
        const double sl0 = (y[1] - y[0]) / dx0, sl1 = (y[2] - y[1]) / dx1;
        /* [1 1 0; dx1 2(dx0+dx1) dx0; 0 1 1] s = [2 sl0; 3(dx0 sl1 + dx1 sl0); 2 sl1] */
        const double b0 = 2 * sl0, b1 = 3 * (dx0 * sl1 + dx1 * sl0), b2 = 2 * sl1;
        /* s0 = b0 - s1, s2 = b2 - s1  =>  dx1 (b0 - s1) + 2 (dx0 + dx1) s1 + dx0 (b2 - s1) = b1 */
        const double s1 = (b1 - dx1 * b0 - dx0 * b2) / (dx0 + dx1);
        s[0] = b0 - s1; s[1] = s1; s[2] = b2 - s1;
        return 0;
    }
    /* tridiagonal system, Thomas algorithm: w[0..n) modified upper diagonal, s = modified rhs */
    double *cp = w;
    {
        const double dx0 = x[1] - x[0], dx1 = x[2] - x[1], d = x[2] - x[0];
        const double sl0 = (y[1] - y[0]) / dx0, sl1 = (y[2] - y[1]) / dx1;
        const double diag = dx1, up = d;
        const double rhs = ((dx0 + 2 * d) * dx1 * sl0 + dx0 * dx0 * sl1) / d;
        cp[0] = up / diag;
        s[0] = rhs / diag;
    }
    for (int i = 1; i < n - 1; ++i) {
        const double dxm = x[i] - x[i - 1], dxp = x[i + 1] - x[i];
        const double slm = (y[i] - y[i - 1]) / dxm, slp = (y[i + 1] - y[i]) / dxp;
        const double lo = dxp, diag = 2 * (dxm + dxp), up = dxm;
        const double rhs = 3 * (dxp * slm + dxm * slp);
        const double den = diag - lo * cp[i - 1];
        cp[i] = up / den;
        s[i] = (rhs - lo * s[i - 1]) / den;
    }
    {
        const int i = n - 1;
        const double dxm = x[i] - x[i - 1], dxmm = x[i - 1] - x[i - 2], d = x[i] - x[i - 2];
        const double slm = (y[i] - y[i - 1]) / dxm, slmm = (y[i - 1] - y[i - 2]) / dxmm;
        const double lo = d, diag = dxmm;
        const double rhs = (dxm * dxm * slmm + (2 * d + dxm) * dxmm * slm) / d;
        const double den = diag - lo * cp[i - 1];
        s[i] = (rhs - lo * s[i - 1]) / den;
    }
    for (int i = n - 2; i >= 0; --i) s[i] -= cp[i] * s[i + 1];
    return 0;
}

static double spline_eval(int n, const double *x, const double *y, const double *s, double t)
{
    int i = 0;                                   /* interval: x[i] <= t < x[i+1], last one closed */
    while (i < n - 2 && t >= x[i + 1]) ++i;
    const double h = x[i + 1] - x[i], slope = (y[i + 1] - y[i]) / h;
    const double tt = (s[i] + s[i + 1] - 2 * slope) / h;
    const double c0 = tt / h, c1 = (slope - s[i]) / h - tt, c2 = s[i], c3 = y[i];
    const double d = t - x[i];
    return c3 + c2 * d + c1 * (d * d) + c0 * (d * d * d);
}

int oracle_initialization(const oracle_problem *p, double t0, double *x0)
{
    const int D = p->D, N = p->N, M = p->M, n = M + 2, DD = D * D;
    if (M < 1) return 1;
    double *x = (double *)malloc(sizeof(double) * n), *y = (double *)malloc(sizeof(double) * n);
    double *s = (double *)malloc(sizeof(double) * n), *w = (double *)malloc(sizeof(double) * 2 * n);
    double *mt0 = (double *)malloc(sizeof(double) * (size_t)N * D);
    double *a0 = x0, *b0 = x0 + (size_t)N * DD;
    int rc = 0;
    x[0] = t0;                                                /* time_x (:86) */
    for (int j = 0; j < M; ++j) x[1 + j] = t0 + (double)p->obs_t[j] * p->dt_model;
    x[n - 1] = t0 + (double)(N - 1) * p->dt_model;
    for (int d = 0; d < D && !rc; ++d) {
        y[0] = p->obs_y[d];                                   /* obs_z (:91 / :104) */
        for (int j = 0; j < M; ++j) y[1 + j] = p->obs_y[(size_t)j * D + d];
        y[n - 1] = p->obs_y[(size_t)(M - 1) * D + d];
        rc = spline_slopes(n, x, y, s, w);
        for (int k = 0; k < N && !rc; ++k)
            mt0[(size_t)k * D + d] = spline_eval(n, x, y, s, t0 + (double)k * p->dt_model);
    }
    if (!rc) {
        if (D == 1) {                                         /* :94-101 */
            for (int k = 0; k < N; ++k) { a0[k] = 0.5 * (p->sigma[0] / 0.25) * 1.0; b0[k] = mt0[k]; }
        } else {                                              /* :103-133 */
            memset(a0, 0, sizeof(double) * (size_t)N * DD);
            for (int k = 0; k < N; ++k)
                for (int d = 0; d < D; ++d) {
                    const double ad = 0.5 * (p->sigma[d] / 0.25);
                    a0[(size_t)k * DD + d * D + d] = ad;
                    const double m0 = mt0[(size_t)k * D + d];
                    b0[(size_t)k * D + d] = (k < N - 1) ? (mt0[(size_t)(k + 1) * D + d] - m0) / p->dt_model + ad * m0 : ad * m0; /* self.dt = model.time_step (:57) */
                }
        }
    }
    free(x); free(y); free(s); free(w); free(mt0);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* observation energy: gaussian_like.py:69-153; jump tables :155-243          */
/* ------------------------------------------------------------------------ */
double oracle_eobs(const oracle_problem *p, const double *mt, const double *st)
{
    const int D = p->D, M = p->M;
    if (D == 1) { /* gauss_1d :69-96 */
        double s = 0.0;
        for (int n = 0; n < M; ++n) {
            const long long t = p->obs_t[n];
            const double y = p->obs_y[n], E2 = mt[t] * mt[t] + st[t];
            s += (y * y) - 2.0 * y * mt[t] + E2;
        }
        return 0.5 * s / p->R[0] + 0.5 * M * (LOG2PI + log(p->R[0]));
    }
    /* gauss_nd :98-153, H = I, R diagonal.  NOTE sn_diag[n]: the covariance
     * diagonal is indexed by the observation ORDINAL n, not by obs_t[n] (F5). */
    double e = 0.0, logdet = 0.0;
    for (int i = 0; i < D; ++i) logdet += log(sqrt(p->R[i]));
    logdet *= 2.0; /* utilities.log_det :104 */
    for (int n = 0; n < M; ++n) {
        const long long t = p->obs_t[n];
        double zz = 0.0, tr = 0.0;
        for (int i = 0; i < D; ++i) {
            const double zi = (p->obs_y[(size_t)n * D + i] - mt[(size_t)t * D + i]) / sqrt(p->R[i]);
            zz += zi * zi;
            tr += (1.0 / p->R[i]) * st[(size_t)n * D * D + (size_t)i * D + i];
        }
        e += zz + tr;
    }
    return 0.5 * (e + M * (D * LOG2PI + logdet));
}
/* jm (N,D), js (N,D,D): zero except at obs_t */
void oracle_eobs_grad(const oracle_problem *p, const double *mt, double *jm, double *js)
{
    const int D = p->D, N = p->N, M = p->M;
    memset(jm, 0, sizeof(double) * (size_t)N * D);
    memset(js, 0, sizeof(double) * (size_t)N * D * D);
    for (int n = 0; n < M; ++n) {
        const long long t = p->obs_t[n];
        for (int i = 0; i < D; ++i) {
            /* :188 / :235   -(y - m)/R ;   :191 / :238   0.5 / R */
            jm[(size_t)t * D + i] = -(p->obs_y[(size_t)n * D + i] - mt[(size_t)t * D + i]) / p->R[i];
            js[(size_t)t * D * D + (size_t)i * D + i] = 0.5 / p->R[i];
        }
    }
}

/* ------------------------------------------------------------------------ */
/* gradient assembly: variational.py:202-334                                  */
/* ------------------------------------------------------------------------ */
void oracle_gradient(const oracle_problem *p, const double *x, const double *mt,
                     const double *st, const double *lam, const double *psi,
                     const double *Ef, const double *Edf, double *grad)
{
    const int D = p->D, N = p->N, DD = D * D;
    const double *A = x, *b = x + (size_t)N * DD;
    double *gA = grad, *gb = grad + (size_t)N * DD;
    double *w = (double *)malloc(sizeof(double) * (3 * DD + 2 * D));
    double *EA = w, *P1 = EA + DD, *P2 = P1 + DD, *db = P2 + DD, *am = db + D;
    for (int k = 0; k < N; ++k) {
        const double *Ak = A + (size_t)k * DD, *Sk = st + (size_t)k * DD, *mk = mt + (size_t)k * D;
        const double *lk = lam + (size_t)k * D, *Pk = psi + (size_t)k * DD;
        /* _dEsde_db :324-334   inv_sigma (-Efx - A m + b) */
        matvec(D, Ak, mk, am);
        for (int i = 0; i < D; ++i)
            db[i] = (1.0 / p->sigma[i]) * (-Ef[(size_t)k * D + i] - am[i] + b[(size_t)k * D + i]);
        /* _dEsde_da :312-322   inv_sigma (Edf + A) S - outer(db, m) */
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j)
                EA[i * D + j] = (1.0 / p->sigma[i]) * (Edf[(size_t)k * DD + i * D + j] + Ak[i * D + j]);
        matmul(D, EA, Sk, P1);
        matmul(D, Pk, Sk, P2);
        /* _grad_at :300-310 and :280, scaled by dt = model.time_step :284-285 */
        for (int i = 0; i < D; ++i) {
            for (int j = 0; j < D; ++j)
                gA[(size_t)k * DD + i * D + j] =
                    p->dt_model * ((P1[i * D + j] - db[i] * mk[j]) - lk[i] * mk[j] - 2.0 * P2[i * D + j]);
            gb[(size_t)k * D + i] = p->dt_model * (db[i] + lk[i]);
        }
    }
    free(w);
}

/* ------------------------------------------------------------------------ */
/* VarGP.free_energy (+ gradient): variational.py:141-289                     */
/* Any out pointer except F may be NULL.  Returns 0 ok, 1 bad argument,       */
/* 2 covariance not positive definite (numpy LinAlgError in the reference).   */
/* ------------------------------------------------------------------------ */
int oracle_eval(const oracle_problem *p, const double *x, double *F, double *parts /* E0,Esde,Eobs */,
                double *grad, double *o_mt, double *o_st, double *o_lam, double *o_psi,
                double *o_Ef, double *o_Edf, double *o_dEm, double *o_dEs)
{
    const size_t D = p->D, N = p->N, DD = D * D;
    const size_t nv = N * D, nm = N * DD;
    double *buf = (double *)malloc(sizeof(double) * (6 * nm + 6 * nv + N));
    if (!buf) return 1;
    double *st = buf, *psi = st + nm, *Edf = psi + nm, *dEs = Edf + nm, *js = dEs + nm,
           *spare = js + nm, *mt = spare + nm, *lam = mt + nv, *Ef = lam + nv, *dEm = Ef + nv,
           *jm = dEm + nv, *spv = jm + nv, *esde_t = spv + nv;
    (void)spare; (void)spv;
    int rc = oracle_fwd(p, x, mt, st);
    double Esde = 0.0, Eobs = 0.0;
    if (!rc) {
        Eobs = oracle_eobs(p, mt, st);
        rc = oracle_energy(p, x, mt, st, &Esde, esde_t, Ef, Edf, dEm, dEs);
    }
    if (!rc) {
        oracle_eobs_grad(p, mt, jm, js);
        rc = oracle_bwd(p, x, dEm, dEs, jm, js, lam, psi);
    }
    if (!rc) {
        *F = p->E0 + Esde + Eobs;
        if (parts) { parts[0] = p->E0; parts[1] = Esde; parts[2] = Eobs; }
        if (grad) oracle_gradient(p, x, mt, st, lam, psi, Ef, Edf, grad);
        if (o_mt) memcpy(o_mt, mt, sizeof(double) * nv);
        if (o_st) memcpy(o_st, st, sizeof(double) * nm);
        if (o_lam) memcpy(o_lam, lam, sizeof(double) * nv);
        if (o_psi) memcpy(o_psi, psi, sizeof(double) * nm);
        if (o_Ef) memcpy(o_Ef, Ef, sizeof(double) * nv);
        if (o_Edf) memcpy(o_Edf, Edf, sizeof(double) * nm);
        if (o_dEm) memcpy(o_dEm, dEm, sizeof(double) * nv);
        if (o_dEs) memcpy(o_dEs, dEs, sizeof(double) * nm);
    }
    free(buf);
    return rc;
}

/* A batch of independent problems sharing one descriptor shape but each with
 * its own x / observations / noise: used by bench.py's cpu_baseline leg with
 * one OpenMP thread per problem.  probs: array of B descriptors. */
int oracle_eval_batch(const oracle_problem *probs, int B, const double *x, long long x_stride,
                      double *F, double *grad, long long g_stride, int threads)
{
    int rc = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    omp_set_max_active_levels(1);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < B; ++i) {
        int e = oracle_eval(&probs[i], x + (size_t)i * x_stride, &F[i], NULL,
                            grad ? grad + (size_t)i * g_stride : NULL, NULL, NULL, NULL, NULL,
                            NULL, NULL, NULL, NULL);
        if (e) {
#pragma omp atomic write
            rc = e;
        }
    }
    return rc;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Data generation.  The standard-normal draws are inputs, in the layout the
 * reference draws them; built with -ffp-contract=off, so every line below is the reference's
 * sequence of IEEE operations.
 *   DoubleWell.make_trajectory         src/dynamics/double_well.py:122-166
 *   OrnsteinUhlenbeck.make_trajectory  src/dynamics/ornstein_uhlenbeck.py:128-161
 *   Lorenz63.make_trajectory / l63     src/dynamics/lorenz_63.py:181-233, :8-37
 *   Lorenz96.make_trajectory / l96     src/dynamics/lorenz_96.py:249-314, :85-101
 * theta: DW [theta], OU [theta, mu], L63 [sigma, rho, beta], L96 [F].  sigma: (D) diagonal.
 * x_init: (D) state at t0 (DW / OU: required; L63 / L96: NULL = 5000 burn-in steps of 1e-3).
 * z: (D, N) draws; path: (N, D).
 * ------------------------------------------------------------------------------------------ */
static void gen_l63(const double *x, const double *u, double *f)
{
    f[0] = u[0] * (x[1] - x[0]);
    f[1] = (u[1] - x[2]) * x[0] - x[1];
    f[2] = x[0] * x[1] - u[2] * x[2];
}

static void gen_l96(const double *x, double u, int D, double *f)
{
    for (int i = 0; i < D; ++i)                       /* (roll(x,-1) - roll(x,+2)) * roll(x,+1) - x + u */
        f[i] = (x[(i + 1) % D] - x[(i + D - 2) % D]) * x[(i + D - 1) % D] - x[i] + u;
}

int oracle_make_trajectory(int model, int N, double dt, const double *theta, const double *sigma,
                           const double *x_init, const double *z, double *path)
{
    if (model == MODEL_DW || model == MODEL_OU) {
        if (!x_init) return 1;
        const double sq = sqrt(sigma[0] * dt);
        double x = x_init[0];
        path[0] = x;
        for (int t = 1; t < N; ++t) {
            const double ek = sq * z[t];
            if (model == MODEL_DW) x = x + 4.0 * x * (theta[0] - x * x) * dt + ek;     /* double_well.py:158-159 */
            else x = x + theta[0] * (theta[1] - x) * dt + ek;                           /* ornstein_uhlenbeck.py:155 */
            path[t] = x;
        }
        return 0;
    }
    const int D = (model == MODEL_L63) ? 3 : 40;
    double x[40], f[40], sq[40];
    for (int i = 0; i < D; ++i) sq[i] = sqrt(sigma[i] * dt);    /* cholesky of a diagonal matrix */
    if (x_init) {
        memcpy(x, x_init, sizeof(double) * D);
    } else {
        for (int i = 0; i < D; ++i) x[i] = (model == MODEL_L63) ? 1.0 : theta[0];
        if (model == MODEL_L96) x[D / 2] += 1.0e-3;
        for (int k = 0; k < 5000; ++k) {
            if (model == MODEL_L63) gen_l63(x, theta, f); else gen_l96(x, theta[0], D, f);
            for (int i = 0; i < D; ++i) x[i] = x[i] + f[i] * 1.0e-3;
        }
    }
    memcpy(path, x, sizeof(double) * D);
    for (int t = 1; t < N; ++t) {
        if (model == MODEL_L63) gen_l63(x, theta, f); else gen_l96(x, theta[0], D, f);
        for (int i = 0; i < D; ++i) x[i] = x[i] + f[i] * dt + sq[i] * z[(size_t)i * N + t];
        memcpy(path + (size_t)t * D, x, sizeof(double) * D);
    }
    return 0;
}

/* StochasticProcess.collect_obs (stochastic_process.py:177-226): xi is (D, M), obs_y is (M, D). */
void oracle_collect_obs(int D, int M, const long long *obs_t, const double *R, const double *path,
                        const double *xi, double *obs_y)
{
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < D; ++i)
            obs_y[(size_t)j * D + i] = path[(size_t)obs_t[j] * D + i] + sqrt(R[i]) * xi[(size_t)i * M + j];
}
