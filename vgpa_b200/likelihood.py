"""
GaussianLikelihood (src/var_bayes/likelihood.py:13-100 + gaussian_like.py:14-243):
holds the observations and evaluates Eobs and its jump tables ON THE GPU
(vgpa_obs_energy).  Identity observation operator and diagonal noise, as the
sim_params JSON schema produces them.
"""
import numpy as np

from . import engine


class GaussianLikelihood(object):
    __slots__ = ("obs_v", "obs_t", "obs_n", "obs_h", "single_dim", "device")

    def __init__(self, values, times, noise, operator=None, single_dim=True, device=0):
        self.obs_t = np.asarray(times)
        self.obs_v = np.asarray(values)
        self.obs_n = np.asarray(noise)
        if operator is None:
            y0 = self.obs_v[0]
            self.obs_h = np.asarray(1) if y0.ndim == 0 else np.eye(y0.size)
        else:
            self.obs_h = np.asarray(operator)
            eye = np.asarray(1) if self.obs_h.ndim == 0 else np.eye(self.obs_h.shape[0])
            if self.obs_h.shape != eye.shape or np.any(self.obs_h != eye):
                raise ValueError(" GaussianLikelihood: the CUDA path supports the identity"
                                 " observation operator only.")
        self.single_dim = single_dim
        self.device = device

    values = property(lambda self: self.obs_v)
    times = property(lambda self: self.obs_t)
    operator = property(lambda self: self.obs_h)

    @property
    def noise(self):
        return self.obs_n

    @noise.setter
    def noise(self, new_value):
        self.obs_n = new_value

    def noise_diag(self):
        r = np.asarray(self.obs_n, dtype=float)
        if r.ndim == 2:
            if np.count_nonzero(r - np.diag(np.diagonal(r))):
                raise ValueError(" GaussianLikelihood: the CUDA path supports a diagonal"
                                 " observation noise only.")
            return np.diagonal(r).copy()
        return np.atleast_1d(r)

    def __call__(self, m, s):
        """Eobs (gaussian_like.py:69-153)."""
        return engine.obs_energy(self.obs_t, self.obs_v, self.noise_diag(), m, s, self.device)[0]

    def gradients(self, m, s=None):
        """dEobs_dm, dEobs_ds, dEobs_dr (gaussian_like.py:155-243).  As in the reference, the 1-D
        dEobs_dr needs the marginal variances s; the n-D one is all zeros of shape (N, M, M)."""
        m = np.asarray(m, dtype=float)
        if s is None:
            s = np.zeros(m.shape + (() if m.ndim == 1 else (m.shape[1],)))
        _, jm, js, dr = engine.obs_energy(self.obs_t, self.obs_v, self.noise_diag(), m, s, self.device,
                                          with_dr=True)
        return jm, js, dr
