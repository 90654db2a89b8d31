"""
PriorKL0 (src/var_bayes/prior_kl0.py:8-92): the KL term at t = 0.  m0 and s0 are
not optimised (variational.py:183-185), so this is ONE scalar per inference problem,
evaluated once on the host when the problem is set up and handed to the CUDA
library as `E0` (include/vgpa_b200.h).  It is not part of the per-iteration path.
"""
import numpy as np


def _chol_inv(x):
    c_inv = np.linalg.solve(np.linalg.cholesky(x), np.eye(x.shape[0]))
    return c_inv.T.dot(c_inv)


def _log_det(x):
    return 2.0 * np.sum(np.log(np.linalg.cholesky(x).diagonal()))


class PriorKL0(object):
    __slots__ = ("mu0", "tau0", "single_dim")

    def __init__(self, mu0, tau0, single_dim=True):
        self.mu0, self.tau0, self.single_dim = mu0, tau0, single_dim

    def __call__(self, m0, s0):
        return self.gauss_1d(m0, s0) if self.single_dim else self.gauss_nd(m0, s0)

    def gauss_1d(self, m0, s0):
        # prior_kl0.py:58-62 (note -log(s0), not -0.5 log(s0))
        z0 = m0 - self.mu0
        return -np.log(s0) - 0.5 * (1.0 - np.log(self.tau0)) + 0.5 / self.tau0 * (z0 ** 2 + s0)

    def gauss_nd(self, m0, s0):
        # prior_kl0.py:78-90: z0.T.dot(z0) is a SCALAR added to every entry of (s0 - tau0)
        z0 = m0 - self.mu0
        return 0.5 * (_log_det(self.tau0.dot(_chol_inv(s0))) +
                      np.sum(np.diag(_chol_inv(self.tau0).dot(z0.T.dot(z0) + s0 - self.tau0))))

    def gradients(self, m0, s0, lam0, psi0):
        """dKL0/dm0, dKL0/ds0 (prior_kl0.py:94-175): host arithmetic on the t = 0 Lagrange multipliers
        (`VarGP.arg_out["lamt"][0]`, `["psit"][0]`); like __call__ it is not on the per-iteration path."""
        z0 = m0 - self.mu0
        if self.single_dim:
            return lam0 + z0 / self.tau0, psi0 + 0.5 * (1.0 / self.tau0 - 1.0 / s0)
        return (lam0 + np.linalg.solve(self.tau0, z0.T).T,
                psi0 + 0.5 * (_chol_inv(self.tau0) - _chol_inv(s0)))
