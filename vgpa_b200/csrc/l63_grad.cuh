// l63_grad.cuh -- one entry of the Lorenz-63 gradient assembly (variational.py:280-334 with the drift
// moments of lorenz_63.py:319-326), written with EXPLICIT rounding (__fma_rn / __dmul_rn / __dadd_rn) so
// that the one-thread-per-problem kernel (small_dim.cu) and the lane-parallel kernel (l63_lanes.cu),
// whose surrounding control flow differs, cannot be contracted differently by the compiler: a problem's
// gradient is the same bits whichever kernel serves it.
#pragma once

namespace vgpa {

struct L63Row {
    double db;       // isg_i (-<f>_i - (A m)_i + b_i)
    double c[3];     // isg_i (<df/dx>_ik + A_ik)
};

// row i of the terms that do not depend on the column: m = m(t), Sx = S[2][0] (row 1) or S[1][0] (row 2),
// Ar = row i of A(t)
__device__ __forceinline__ L63Row l63_row_terms(int i, double vS, double vR, double vB, const double (&m)[3], double Sx,
                                                const double (&Ar)[3], double bt, double isg)
{
    double Ef, Ed[3];
    if (i == 0) {
        Ef = __dmul_rn(vS, __dsub_rn(m[1], m[0]));
        Ed[0] = -vS; Ed[1] = vS; Ed[2] = 0.0;
    } else if (i == 1) {
        Ef = __fma_rn(-m[0], m[2], __dsub_rn(__fma_rn(vR, m[0], -m[1]), Sx));
        Ed[0] = __dsub_rn(vR, m[2]); Ed[1] = -1.0; Ed[2] = -m[0];
    } else {
        Ef = __fma_rn(-vB, m[2], __fma_rn(m[0], m[1], Sx));
        Ed[0] = m[1]; Ed[1] = m[0]; Ed[2] = -vB;
    }
    const double am = __fma_rn(Ar[2], m[2], __fma_rn(Ar[1], m[1], __dmul_rn(Ar[0], m[0])));
    L63Row r;
    r.db = __dmul_rn(isg, __dadd_rn(__dsub_rn(-Ef, am), bt));
#pragma unroll
    for (int k = 0; k < 3; ++k) r.c[k] = __dmul_rn(isg, __dadd_rn(Ed[k], Ar[k]));
    return r;
}
// dL/dA[t][i][j]: Sc = column j of S(t), Pr = row i of Psi(t), mj = m_j(t), lam = lam_i(t)
__device__ __forceinline__ double l63_grad_a(const L63Row& r, const double (&Sc)[3], const double (&Pr)[3], double mj,
                                             double lam, double dtm)
{
    const double p1 = __fma_rn(r.c[2], Sc[2], __fma_rn(r.c[1], Sc[1], __dmul_rn(r.c[0], Sc[0])));
    const double p2 = __fma_rn(Pr[2], Sc[2], __fma_rn(Pr[1], Sc[1], __dmul_rn(Pr[0], Sc[0])));
    double t = __fma_rn(-r.db, mj, p1);
    t = __fma_rn(-lam, mj, t);
    t = __fma_rn(-2.0, p2, t);
    return __dmul_rn(dtm, t);
}
// dL/db[t][i]
__device__ __forceinline__ double l63_grad_b(const L63Row& r, double lam, double dtm) { return __dmul_rn(dtm, __dadd_rn(r.db, lam)); }

}  // namespace vgpa
