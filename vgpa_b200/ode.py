"""
FwdOde / BwdOde: the reference's two sweep facades (src/var_bayes/fwd_ode.py:14-65,
src/var_bayes/bwd_ode.py:14-65) over the CUDA sweeps (vgpa_solve_fwd / vgpa_solve_bwd).
Same constructor `(dt, method, single_dim)`, same call signatures, same ValueErrors.
"""
import numpy as np

from . import engine
from ._lib import METHODS


class _Ode(object):
    __slots__ = ("dt", "method", "single_dim", "device")

    def __init__(self, dt, method, single_dim=True, device=0):
        if dt <= 0.0:
            raise ValueError(f" {self.__class__.__name__}:"
                             f" Discrete time step should be strictly positive -> {dt}.")
        if str(method).lower() not in METHODS:
            raise ValueError(f" {self.__class__.__name__}: Integration method is unknown -> {method}.")
        self.dt, self.method, self.single_dim, self.device = dt, method, single_dim, device

    def __str__(self):
        return f" {self.__class__.__name__} Id({id(self)}): dt={self.dt}, method={self.method}"


def _diag(sigma):
    s = np.asarray(sigma, dtype=float)
    return np.diagonal(s).copy() if s.ndim == 2 else np.atleast_1d(s)


class FwdOde(_Ode):
    def __call__(self, at, bt, m0, s0, sigma):
        """-> mt, st (marginal moments), fwd_ode.py:45."""
        return engine.solve_fwd(self.method, at, bt, m0, s0, _diag(sigma), self.dt, self.device)


class BwdOde(_Ode):
    def __call__(self, at, dEsde_dm, dEsde_ds, dEobs_dm, dEobs_ds):
        """-> lam, psi (Lagrange multipliers), bwd_ode.py:45."""
        return engine.solve_bwd(self.method, at, dEsde_dm, dEsde_ds, dEobs_dm, dEobs_ds, self.dt,
                                self.device)
