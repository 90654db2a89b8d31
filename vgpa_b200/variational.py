"""
VarGP: drop-in for the reference's src/var_bayes/variational.py:6-336.

Same constructor, same `initialization()`, `free_energy(x)`, `gradient(x, eval_fun)`
and `arg_out`, and the same calling protocol the SCG optimiser relies on
(optim_scg.py:99-100,167,189,234-235).  free_energy and gradient are computed by
ONE batched CUDA evaluation (B = 1 here) through the C ABI; the gradient of the
last x is cached so that `f(x)` followed by `df(x)` costs one evaluation instead
of the reference's two.  There is no CPU path.
"""
import ctypes as C
import weakref

import numpy as np
from scipy.interpolate import CubicSpline

from ._lib import PinnedArray, lib
from .engine import BatchEvaluator

class _GradientPool:
    """Page-locked buffers for the gradient, handed to the caller WITHOUT a copy.

    `VarGP.gradient` must return an array the caller owns (optim_scg.py keeps the last two
    gradients alive).  Copying 13 MB out of a staging buffer and first-touching a fresh
    allocation costs about 2 ms per call at the L96 shape, so the device-to-host copy lands
    directly in the array that is returned.  Ownership is explicit: every array handed out is a
    fresh ndarray object over one page-locked buffer with a `weakref.finalize` on it; views keep
    that root array alive (their `.base`), so the finalizer runs exactly when the caller holds
    nothing of the buffer any more, and only then does the buffer go back to the free list (or,
    after `close()`, to `cudaFreeHost`)."""

    def __init__(self, n, max_buffers=8):
        self.n, self.max_buffers = int(n), int(max_buffers)
        self._free, self._allocated, self._closed = [], 0, False

    def take(self):
        if self._free:
            ptr = self._free.pop()
        elif self._allocated < self.max_buffers:
            ptr = lib.vgpa_host_alloc(max(self.n, 1) * 8)
            if not ptr:
                return None
            self._allocated += 1
        else:
            return None                                   # caller hoards gradients: fall back to copies
        buf = (C.c_double * max(self.n, 1)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=np.float64, count=self.n)
        fin = weakref.finalize(arr, _GradientPool._release, weakref.ref(self), ptr)
        fin.atexit = False                                # at interpreter exit the driver reclaims it
        return arr

    @staticmethod
    def _release(pool_ref, ptr):
        pool = pool_ref()
        if pool is None or pool._closed:
            lib.vgpa_host_free(ptr)
        else:
            pool._free.append(ptr)

    @property
    def outstanding(self):
        """Buffers currently held by callers (or by the VarGP cache)."""
        return self._allocated - len(self._free)

    def close(self):
        self._closed = True                               # buffers still held outside free themselves later
        while self._free:
            lib.vgpa_host_free(self._free.pop())
            self._allocated -= 1

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _diag(a, what):
    a = np.asarray(a, dtype=float)
    if a.ndim == 2:
        if np.count_nonzero(a - np.diag(np.diagonal(a))):
            raise ValueError(f" VarGP: the CUDA path supports a diagonal {what} only.")
        return np.diagonal(a).copy()
    return np.atleast_1d(a)


# the reference's dynamics classes carry no model key: recognise them by class name, so that the
# reference's own objects can be handed to this VarGP unmodified
_MODEL_KEY_BY_CLASS = {"DoubleWell": "DW", "OrnsteinUhlenbeck": "OU", "Lorenz63": "L63", "Lorenz96": "L96"}


def _model_key(model):
    key = getattr(model, "model_key", None)
    if key is None:
        for cls in type(model).__mro__:
            if cls.__name__ in _MODEL_KEY_BY_CLASS:
                return _MODEL_KEY_BY_CLASS[cls.__name__]
        raise ValueError(f" VarGP: the CUDA path has no kernel for dynamics {type(model).__name__!r} "
                         "(DoubleWell, OrnsteinUhlenbeck, Lorenz63, Lorenz96).")
    return key


class VarGP(object):

    def __init__(self, model, m0, s0, fwd_ode, bwd_ode, likelihood, kl0, obs_y, obs_t, device=0):
        self.model = model
        self.fwd_ode, self.bwd_ode = fwd_ode, bwd_ode
        self.kl0, self.likelihood = kl0, likelihood
        self.obs_y, self.obs_t = obs_y, obs_t
        self.dt = self.model.time_step                           # variational.py:57
        if self.model.single_dim:
            self.dim_n, self.dim_d = self.model.sample_path.size, 1
        else:
            self.dim_n, self.dim_d = self.model.sample_path.shape
        self.dim_tot = self.dim_n * self.dim_d * self.dim_d
        self.output = {"m0": m0, "s0": s0}
        if str(fwd_ode.method).lower() != str(bwd_ode.method).lower() or fwd_ode.dt != bwd_ode.dt:
            raise ValueError(" VarGP: forward and backward sweeps must share method and step.")
        self._device = device
        self._ev_obj = None
        self._x_cached = None
        self._x_probe = None
        self._f_cached = None
        self._g_cached = None
        self._full_for = None
        self._pin = None
        self._g_owned = False
        self.n_eval = 0

    @property
    def _ev(self):
        """The CUDA handle, created on first use (so that set-up code runs without a GPU)."""
        if self._ev_obj is None:
            D = self.dim_d
            m0, s0 = self.output["m0"], self.output["s0"]
            self._ev_obj = BatchEvaluator(
                model=_model_key(self.model), method=self.fwd_ode.method, N=self.dim_n, dt=self.fwd_ode.dt,
                theta=np.atleast_1d(np.asarray(self.model.theta, dtype=float)),
                sigma=_diag(self.model.sigma, "system noise"),
                R=_diag(self.likelihood.noise, "observation noise"),
                obs_t=np.asarray(self.obs_t, dtype=np.int64),
                obs_y=np.asarray(self.obs_y, dtype=float).reshape(-1, D),
                m0=np.atleast_1d(np.asarray(m0, dtype=float)),
                s0=np.asarray(s0, dtype=float).reshape(D, D),
                E0=float(np.asarray(self.kl0(m0, s0))),              # constant: variational.py:183-185
                B=1, dt_model=float(self.dt), device=self._device)
        return self._ev_obj

    # -- variational.py:73-139 (host-side set-up, not on the per-iteration path) ------
    def initialization(self):
        time_window = self.model.time_window
        time_x = [time_window[0], *time_window[self.obs_t], time_window[-1]]
        if self.model.single_dim:
            obs_z = np.hstack((self.obs_y[0], self.obs_y, self.obs_y[-1]))
            a0 = 0.5 * (self.model.sigma / 0.25) * np.ones(self.dim_n)
            b0 = CubicSpline(time_x, obs_z)(time_window)
        else:
            obs_z = np.vstack((self.obs_y[0], self.obs_y, self.obs_y[-1]))
            mt0 = CubicSpline(time_x, obs_z)(time_window)
            a0 = np.zeros((self.dim_n, self.dim_d, self.dim_d))
            b0 = np.zeros((self.dim_n, self.dim_d))
            s0 = 0.25 * np.eye(self.dim_d)
            dmt0 = np.diff(mt0, axis=0) / self.dt
            diag_k = np.diag(self.model.sigma.diagonal() / s0.diagonal())
            for k in range(self.dim_n - 1):
                a0[k] = 0.5 * diag_k
                b0[k] = dmt0[k] + a0[k].diagonal() * mt0[k]
            a0[-1] = 0.5 * diag_k
            b0[-1] = a0[-1].diagonal() * mt0[-1]
        return np.concatenate((a0.ravel(), b0.ravel()))

    def initialization_gpu(self):
        """The same x0 computed by the CUDA library (vgpa_initialization_host): the single-problem
        view of the batched on-device initialisation that ensembles use (BatchEvaluator.initialization /
        .initialization_device); agrees with initialization() to rounding (tests/test_gpu_init.py)."""
        return self._ev.initialization(float(self.model.time_window[0]))[0]

    # -- the hot path -------------------------------------------------------------------
    def _evaluate(self, x):
        """One CUDA evaluation of F and the gradient at x.  x is staged ONCE into a page-locked buffer
        that is also the cache of the last point, and the gradient lands in a second page-locked
        buffer, so both PCIe copies are plain DMA (no driver-side staging of pageable memory, no
        first-touch page faults of a fresh 13 MB array per call)."""
        ev = self._ev
        if self._pin is None:
            self._pin = (PinnedArray((ev.n_x,)), _GradientPool(ev.n_x), np.empty(1))
        px, pool, F = self._pin[0].array, self._pin[1], self._pin[2]
        xa = np.asarray(x).reshape(-1)
        if xa.size != ev.n_x:
            raise ValueError(f"x: expected {ev.n_x} values, got {xa.size}")
        self._x_cached = None                   # px is about to change
        self._g_cached = None                   # ... and our reference must not keep a buffer busy
        g = pool.take()
        self._g_owned = g is not None           # the caller may receive this very array, once
        if g is None:
            g = np.empty(ev.n_x)
        if xa.dtype == np.float64 and xa.flags.c_contiguous and xa.nbytes >= (4 << 20):
            lib.vgpa_host_copy(px.ctypes.data, xa.ctypes.data, xa.nbytes, 4)      # 13 MB at L96: four threads
        else:
            np.copyto(px, xa)
        ev.eval(px, want_grad=True, F_out=F, G_out=g.reshape(1, -1))
        self.n_eval += 1
        self._x_cached = px
        self._x_probe = px[::self._probe_step(px.size)].copy()
        self._f_cached = float(F[0])
        self._g_cached = g
        self._full_for = None

    @staticmethod
    def _probe_step(n):
        return max(1, n // 2048)

    def _is_cached(self, x):
        """Was the last evaluation at this x?  (The reference does not check at all: it recomputes in
        free_energy(x) and its gradient(x) trusts the caller, variational.py:141-226.)  A strided probe
        of ~2000 entries rejects a new point without touching the 13 MB array; a point that passes the
        probe is ALWAYS compared in full against the page-locked copy of the last x (multi-threaded
        memcmp, ~0.4 ms at the L96 shape), so an array mutated in place between calls -- coordinate-wise
        finite differences, sparse updates -- is never served from the cache."""
        if self._x_cached is None:
            return False
        xa = np.asarray(x).reshape(-1)
        if xa.size != self._x_cached.size:
            return False
        if not np.array_equal(xa[::self._probe_step(xa.size)], self._x_probe):
            return False
        if xa.dtype == np.float64 and xa.flags.c_contiguous:
            return bool(lib.vgpa_host_equal(xa.ctypes.data, self._x_cached.ctypes.data, xa.nbytes, 4))
        return np.array_equal(xa, self._x_cached)

    def free_energy(self, x):
        """E0 + Esde + Eobs as a Python float (variational.py:141-200)."""
        if not self._is_cached(x):
            self._evaluate(x)
        return self._f_cached

    def gradient(self, x, eval_fun=False):
        """[dL/dA | dL/db] (variational.py:202-289).  `eval_fun=True` asks for a
        consistent state at a new x; the cache makes that automatic."""
        if not self._is_cached(x) or self._g_cached is None:
            self._evaluate(x)
        if self._g_owned:
            # hand the page-locked buffer over: from here on it is the caller's array (it may even be
            # modified in place), so this object forgets it; asking again at the same x re-evaluates
            g, self._g_cached, self._g_owned = self._g_cached, None, False
            return g
        return self._g_cached.copy()            # pool exhausted: private buffer, copies out

    @property
    def arg_out(self):
        """m0, s0, mt, st, Efx, Edf, lamt, psit of the last evaluated x (variational.py:292)."""
        if self._x_cached is not None and self._full_for is None:
            full = self._ev.eval_full(self._x_cached)
            for k in ("mt", "st", "Efx", "Edf", "lamt", "psit"):
                self.output[k] = full[k]
            self._full_for = True
        return self.output

    def close(self):
        if self._ev_obj is not None:
            self._ev_obj.close()
            self._ev_obj = None
        if self._pin is not None:
            self._x_cached = self._g_cached = None
            self._pin[0].free()
            self._pin[1].close()
            self._pin = None
