"""
ctypes front-end of oracle/vgpa_oracle.c (the CPU restatement of the reference
path).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "libvgpa_oracle.so"

MODELS = {"DW": 0, "OU": 1, "L63": 2, "L96": 3}
METHODS = {"euler": 0, "heun": 1, "rk2": 2, "rk4": 3}

_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_longlong)


class _CProblem(C.Structure):
    _fields_ = [("model", C.c_int), ("method", C.c_int),
                ("D", C.c_int), ("N", C.c_int), ("M", C.c_int),
                ("dt", C.c_double), ("dt_model", C.c_double),
                ("theta", _dp), ("sigma", _dp), ("R", _dp),
                ("obs_t", _lp), ("obs_y", _dp), ("m0", _dp), ("s0", _dp),
                ("E0", C.c_double)]


def build_oracle(force=False):
    """gcc -O2 -fopenmp -shared; no fast-math, no FMA contraction."""
    src = HERE / "vgpa_oracle.c"
    if (not force and ORACLE_SO.exists()
            and ORACLE_SO.stat().st_mtime >= src.stat().st_mtime):
        return ORACLE_SO
    cmd = ["gcc", "-O2", "-std=c11", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared",
           "-o", str(ORACLE_SO), str(src), "-lm"]
    subprocess.run(cmd, check=True)
    return ORACLE_SO


def prior_kl0(m0, s0, mu0, tau0, single_dim):
    """Restates PriorKL0.gauss_1d / gauss_nd (src/var_bayes/prior_kl0.py:46-92)
    for diagonal tau0/s0 handling by plain numpy.  Note the two quirks kept:
    the 1-D form has -log(s0) (not -0.5 log s0) and the n-D form adds the
    SCALAR z0.z0 to every entry of (s0 - tau0) before the trace."""
    if single_dim:
        m0 = float(np.asarray(m0).ravel()[0]); s0 = float(np.asarray(s0).ravel()[0])
        mu0 = float(np.asarray(mu0).ravel()[0]); tau0 = float(np.asarray(tau0).ravel()[0])
        z0 = m0 - mu0
        return -np.log(s0) - 0.5 * (1.0 - np.log(tau0)) + 0.5 / tau0 * (z0 ** 2 + s0)
    m0 = np.asarray(m0, float); s0 = np.asarray(s0, float)
    mu0 = np.asarray(mu0, float); tau0 = np.asarray(tau0, float)

    def chol_inv(x):  # utilities.py:203-217
        c_inv = np.linalg.solve(np.linalg.cholesky(x), np.eye(x.shape[0]))
        return c_inv.T.dot(c_inv)

    def log_det(x):  # utilities.py:68-105
        return 2.0 * np.sum(np.log(np.linalg.cholesky(x).diagonal()))

    inv_tau0 = chol_inv(tau0)
    inv_s0 = chol_inv(s0)
    z0 = m0 - mu0
    return float(0.5 * (log_det(tau0.dot(inv_s0)) +
                        np.sum(np.diag(inv_tau0.dot(z0.T.dot(z0) + s0 - tau0)))))


@dataclass
class Problem:
    """One inference problem, in the reference's own terms."""
    model: str
    method: str
    D: int
    N: int
    dt: float
    theta: np.ndarray
    sigma: np.ndarray          # (D,) diagonal of the system noise
    R: np.ndarray              # (D,) diagonal of the observation noise
    obs_t: np.ndarray          # (M,) int64
    obs_y: np.ndarray          # (M, D)
    m0: np.ndarray             # (D,)
    s0: np.ndarray             # (D, D)
    E0: float = 0.0
    dt_model: float = None
    _keep: list = field(default_factory=list, repr=False)

    @classmethod
    def from_golden(cls, g):
        """Build from a tests/golden/*.npz record."""
        D, N = int(g["D"]), int(g["N"])
        tk = np.arange(0.0, float(g["tf"]) + float(g["dt"]), float(g["dt"]))
        assert tk.size == N
        E0 = prior_kl0(g["m0"], g["s0"], g["mu0"], g["tau0"], D == 1)
        return cls(model=str(g["model"]), method=str(g["method"]), D=D, N=N, dt=float(g["dt"]),
                   theta=g["theta"], sigma=g["sigma"], R=g["R"], obs_t=g["obs_t"],
                   obs_y=g["obs_y"], m0=g["m0"], s0=g["s0"], E0=float(E0),
                   dt_model=float(np.abs(tk[1] - tk[0])))

    def c_struct(self):
        def arr(a, dt=np.float64):
            a = np.ascontiguousarray(np.asarray(a, dtype=dt).ravel())
            self._keep.append(a)
            return a
        th, sg, R = arr(np.atleast_1d(self.theta)), arr(np.atleast_1d(self.sigma)), arr(np.atleast_1d(self.R))
        ot, oy = arr(self.obs_t, np.int64), arr(self.obs_y)
        m0, s0 = arr(self.m0), arr(self.s0)
        assert sg.size == self.D and R.size == self.D and m0.size == self.D and s0.size == self.D ** 2
        assert oy.size == ot.size * self.D
        return _CProblem(MODELS[self.model.upper()], METHODS[self.method.lower()], self.D, self.N,
                         ot.size, self.dt, self.dt if self.dt_model is None else self.dt_model,
                         th.ctypes.data_as(_dp), sg.ctypes.data_as(_dp), R.ctypes.data_as(_dp),
                         ot.ctypes.data_as(_lp), oy.ctypes.data_as(_dp), m0.ctypes.data_as(_dp),
                         s0.ctypes.data_as(_dp), self.E0)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


class Oracle:
    def __init__(self):
        build_oracle()
        self.lib = C.CDLL(str(ORACLE_SO))
        L = self.lib
        L.oracle_eval.restype = C.c_int
        L.oracle_eval.argtypes = [C.POINTER(_CProblem)] + [_dp] * 12
        L.oracle_fwd.restype = C.c_int
        L.oracle_fwd.argtypes = [C.POINTER(_CProblem), _dp, _dp, _dp]
        L.oracle_bwd.restype = C.c_int
        L.oracle_bwd.argtypes = [C.POINTER(_CProblem)] + [_dp] * 7
        L.oracle_energy.restype = C.c_int
        L.oracle_energy.argtypes = [C.POINTER(_CProblem)] + [_dp] * 9
        L.oracle_initialization.restype = C.c_int
        L.oracle_initialization.argtypes = [C.POINTER(_CProblem), C.c_double, _dp]
        L.oracle_energy_hyper.restype = C.c_int
        L.oracle_energy_hyper.argtypes = [C.POINTER(_CProblem)] + [_dp] * 5
        L.oracle_eobs_dr.restype = None
        L.oracle_eobs_dr.argtypes = [C.POINTER(_CProblem), _dp, _dp, _dp]
        L.oracle_eobs.restype = C.c_double
        L.oracle_eobs.argtypes = [C.POINTER(_CProblem), _dp, _dp]
        L.oracle_eobs_grad.restype = None
        L.oracle_eobs_grad.argtypes = [C.POINTER(_CProblem), _dp, _dp, _dp]
        L.oracle_eval_batch.restype = C.c_int
        L.oracle_eval_batch.argtypes = [C.POINTER(_CProblem), C.c_int, _dp, C.c_longlong, _dp, _dp,
                                        C.c_longlong, C.c_int]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_make_trajectory.restype = C.c_int
        L.oracle_make_trajectory.argtypes = [C.c_int, C.c_int, C.c_double] + [_dp] * 5
        L.oracle_collect_obs.restype = None
        L.oracle_collect_obs.argtypes = [C.c_int, C.c_int, _lp, _dp, _dp, _dp, _dp]

    @staticmethod
    def _raise(rc):
        if rc == 2:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        if rc:
            raise ValueError(f"oracle: invalid argument (rc={rc})")

    def eval(self, prob, x, want_grad=True, full=False):
        """F (and grad); full=True also returns every intermediate as a dict."""
        D, N = prob.D, prob.N
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == N * D * (D + 1)
        cp = prob.c_struct()
        F = np.zeros(1); parts = np.zeros(3)
        grad = np.empty(x.size) if want_grad else None
        outs = {}
        if full:
            for k, n in (("mt", N * D), ("st", N * D * D), ("lamt", N * D), ("psit", N * D * D),
                         ("Efx", N * D), ("Edf", N * D * D), ("dEsde_dm", N * D),
                         ("dEsde_ds", N * D * D)):
                outs[k] = np.empty(n)
        rc = self.lib.oracle_eval(C.byref(cp), _p(x), _p(F), _p(parts), _p(grad),
                                  *[_p(outs.get(k)) for k in ("mt", "st", "lamt", "psit", "Efx",
                                                              "Edf", "dEsde_dm", "dEsde_ds")])
        self._raise(rc)
        if not full:
            return float(F[0]), grad
        shp_v = (N,) if D == 1 else (N, D)
        shp_m = (N,) if D == 1 else (N, D, D)
        for k in outs:
            outs[k] = outs[k].reshape(shp_m if k in ("st", "psit", "Edf", "dEsde_ds") else shp_v)
        outs.update(F=float(F[0]), E0=parts[0], Esde=parts[1], Eobs=parts[2], grad=grad)
        return outs

    def initialization(self, prob, t0=0.0):
        """VarGP.initialization: x0 = [A0 | b0]."""
        x0 = np.zeros(prob.N * prob.D * (prob.D + 1))
        cp = prob.c_struct()
        self._raise(self.lib.oracle_initialization(C.byref(cp), float(t0), _p(x0)))
        return x0

    def energy_hyper(self, prob, x, mt, st):
        """dEsde_dtheta, dEsde_dsigma of model.energy (reference shapes)."""
        D, N = prob.D, prob.N
        nth = {"DW": 1, "OU": 1, "L63": 3, "L96": D}[prob.model.upper()]
        x, mt, st = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, mt, st))
        dth, dsig = np.zeros(nth), np.zeros(1 if D == 1 else D * D)
        cp = prob.c_struct()
        self._raise(self.lib.oracle_energy_hyper(C.byref(cp), _p(x), _p(mt), _p(st), _p(dth), _p(dsig)))
        if D == 1:
            return float(dth[0]), float(dsig[0])
        return dth, dsig.reshape(D, D)

    def eobs_dr(self, prob, mt, st):
        """dEobs_dr of GaussianLikelihood.gradients."""
        D, N, M = prob.D, prob.N, len(np.atleast_1d(prob.obs_t))
        mt, st = (np.ascontiguousarray(a, dtype=np.float64) for a in (mt, st))
        dr = np.zeros(N if D == 1 else N * M * M)
        cp = prob.c_struct()
        self.lib.oracle_eobs_dr(C.byref(cp), _p(mt), _p(st), _p(dr))
        return dr if D == 1 else dr.reshape(N, M, M)

    def eval_batch(self, probs, X, want_grad=True, threads=0):
        """X: (B, n) one evaluation point per problem."""
        B = len(probs)
        X = np.ascontiguousarray(X, dtype=np.float64)
        arr = (_CProblem * B)(*[p.c_struct() for p in probs])
        F = np.zeros(B)
        G = np.empty_like(X) if want_grad else None
        rc = self.lib.oracle_eval_batch(arr, B, _p(X), X.shape[1], _p(F), _p(G),
                                        X.shape[1], threads)
        self._raise(rc)
        return F, G

    def make_trajectory(self, model, N, dt, theta, sigma_diag, z, x_init=None):
        """<Model>.make_trajectory given its standard-normal draws z ((D, N); (N,) for D = 1)."""
        key = model.upper()
        D = {"DW": 1, "OU": 1, "L63": 3, "L96": 40}[key]
        a = [np.ascontiguousarray(np.atleast_1d(v), dtype=np.float64) for v in (theta, sigma_diag, z)]
        xi = None if x_init is None else np.ascontiguousarray(np.atleast_1d(x_init), dtype=np.float64)
        assert a[2].size == N * D and a[1].size == D
        path = np.empty(N * D)
        self._raise(self.lib.oracle_make_trajectory(MODELS[key], int(N), float(dt), _p(a[0]), _p(a[1]), _p(xi),
                                                    _p(a[2]), _p(path)))
        return path if D == 1 else path.reshape(N, D)

    def collect_obs(self, path, obs_t, R_diag, xi):
        """StochasticProcess.collect_obs given its draws xi ((D, M); (M,) for D = 1)."""
        R = np.ascontiguousarray(np.atleast_1d(R_diag), dtype=np.float64)
        D = R.size
        ot = np.ascontiguousarray(np.asarray(obs_t, dtype=np.int64).ravel())
        path, xi = (np.ascontiguousarray(v, dtype=np.float64) for v in (path, xi))
        assert xi.size == D * ot.size and path.size % D == 0
        out = np.empty(ot.size * D)
        self.lib.oracle_collect_obs(D, ot.size, ot.ctypes.data_as(_lp), _p(R), _p(path), _p(xi), _p(out))
        return out if D == 1 else out.reshape(ot.size, D)

    def num_threads(self):
        return int(self.lib.oracle_num_threads())
