"""
BatchedSCG: the reference's scaled-conjugate-gradient optimiser
(src/numerics/optim_scg.py:75-285) run for B independent problems AT ONCE with the
variational parameters, gradients and search directions resident in HBM.

SURVEY.md 8(f) item 1: the host SCG costs a 13 MB host<->device round trip of x and
grad F per evaluation (L96), which caps the end-to-end rate at PCIe speed.  Here the
per-problem control flow is vectorised over B on the host (B scalars per quantity),
while every length-n operation -- the two free-energy/gradient evaluations of an
iteration (vgpa_eval_device) and the optimiser's dot products, AXPYs and direction
updates (vgpa_bdot / baxpy / bdir / bcopy / bstats) -- runs on the device.  PyTorch
only owns the device buffers.

Per problem the arithmetic is the reference's, so each problem's `fx` trace follows
the single-problem SCG (tests/test_gpu_batched_scg.py: 1e-6 on the common prefix).
"""
import numpy as np

from ._lib import lib, raise_for


class BatchedSCG:
    SIGMA0 = 1.0e-3
    BETA_MIN, BETA_MAX = 1.0e-15, 1.0e+100

    def __init__(self, evaluator, options=None):
        """evaluator: a vgpa_b200.BatchEvaluator (B problems, n_x parameters each)."""
        opts = options or {}
        self.ev = evaluator
        self.nit = opts.get("max_it", 150)
        self.x_tol = opts.get("x_tol", 1.0e-6)
        self.f_tol = opts.get("f_tol", 1.0e-8)
        self.display = opts.get("display", False)
        self.stats = None

    # -- device helpers ----------------------------------------------------------
    def _setup(self, X0):
        import torch
        self.torch = torch
        ev = self.ev
        B, n = ev.B, ev.n_x
        dev = torch.device("cuda", ev.device)
        X = torch.as_tensor(X0, dtype=torch.float64)
        if X.dim() == 1:
            X = X.unsqueeze(0).expand(B, n)
        self.X = X.to(dev).contiguous().clone()
        z = lambda: torch.empty((B, n), dtype=torch.float64, device=dev)
        self.XT, self.Gn, self.Go, self.GT, self.Dd = z(), z(), z(), z(), z()
        self.Fd = torch.empty(B, dtype=torch.float64, device=dev)
        self.sc = torch.zeros(3 * B, dtype=torch.float64, device=dev)       # reduction outputs
        self.coef = torch.empty(B, dtype=torch.float64, device=dev)         # per-problem scalars in
        self.mask = torch.empty(B, dtype=torch.int32, device=dev)
        self.act = torch.ones(B, dtype=torch.int32, device=dev)              # active set of the evaluations
        self.actv = torch.ones(B, dtype=torch.int32, device=dev)             # ... of the vector kernels
        self._actv_ptr = None                                                # None = every row
        self.stream = torch.cuda.current_stream(dev).cuda_stream
        self.B, self.n = B, n

    def _eval(self, X, G, who=None):
        """F (host, B) and grad (device) at the rows of X.  `who` (B booleans): the problems that need
        this evaluation; the kernels skip the others (vgpa_set_active), whose F entries and gradient
        rows keep their previous content -- every use below is masked by the same condition.  A batch
        thus stops paying for problems that have converged."""
        if who is not None and not who.all():
            self.act.copy_(self.torch.from_numpy(np.ascontiguousarray(who.astype(np.int32))))
            self.ev.set_active(self.act.data_ptr())
        else:
            self.ev.set_active(None)
        try:
            self.ev.eval_device(X.data_ptr(), self.n, self.Fd.data_ptr(), G.data_ptr(), self.n, self.stream)
            self.ev.sync()
        finally:
            self.ev.set_active(None)
        return self.Fd.cpu().numpy()

    def _rows(self, active):
        """Rows the vector kernels below work on: the problems still being optimised (a superset is
        harmless: every use of their results is masked on the host)."""
        if active.all():
            self._actv_ptr = None
        else:
            self.actv.copy_(self.torch.from_numpy(np.ascontiguousarray(active.astype(np.int32))))
            self._actv_ptr = self.actv.data_ptr()

    def _dot(self, x, y, z=None):
        raise_for(lib.vgpa_bdot(self.B, self.n, x.data_ptr(), y.data_ptr(), z.data_ptr() if z is not None else None,
                                self.n, self.sc.data_ptr(), self._actv_ptr, self.stream))
        r = self.sc.cpu().numpy().reshape(3, self.B)
        return r[0].copy(), r[1].copy(), r[2].copy()

    def _axpy(self, a, x, y, out):
        self.coef.copy_(self.torch.from_numpy(np.ascontiguousarray(a)))
        raise_for(lib.vgpa_baxpy(self.B, self.n, self.coef.data_ptr(), x.data_ptr(), y.data_ptr(), out.data_ptr(),
                                 self.n, self._actv_ptr, self.stream))

    def _copy_where(self, m, src, dst):
        self.mask.copy_(self.torch.from_numpy(np.ascontiguousarray(m.astype(np.int32))))
        raise_for(lib.vgpa_bcopy(self.B, self.n, self.mask.data_ptr(), src.data_ptr(), dst.data_ptr(), self.n,
                                 self.stream))

    def _direction(self, mode, gamma):
        self.mask.copy_(self.torch.from_numpy(np.ascontiguousarray(mode.astype(np.int32))))
        self.coef.copy_(self.torch.from_numpy(np.ascontiguousarray(gamma)))
        raise_for(lib.vgpa_bdir(self.B, self.n, self.mask.data_ptr(), self.coef.data_ptr(), self.Dd.data_ptr(),
                                self.Gn.data_ptr(), self.n, self.stream))

    def _maxabs_sumabs(self, x):
        raise_for(lib.vgpa_bstats(self.B, self.n, x.data_ptr(), self.n, self.sc.data_ptr(), self._actv_ptr, self.stream))
        r = self.sc.cpu().numpy()[:2 * self.B].reshape(2, self.B)
        return r[0].copy(), r[1].copy()

    # -- the optimiser -------------------------------------------------------------
    def __call__(self, X0):
        """Returns (X (B, n) device tensor, fx (B,) numpy).  self.stats holds the per-problem
        traces: fx, dfx, beta of shape (max_it, B), MaxIt (B,), f_eval, df_eval (B,)."""
        self._setup(X0)
        B, nit = self.B, self.nit
        eps = np.finfo(float).eps
        st = {"MaxIt": np.full(B, nit), "fx": np.zeros((nit, B)), "dfx": np.zeros((nit, B)),
              "beta": np.zeros((nit, B)), "f_eval": np.zeros(B), "df_eval": np.zeros(B), "evaluations": 0}
        self.stats = st
        f_now = self._eval(self.X, self.Gn).copy()
        st["evaluations"] += 1
        st["f_eval"] += 1
        st["df_eval"] += 1
        f_old = f_now.copy()
        self.Go.copy_(self.Gn)
        self._direction(np.full(B, 2), np.zeros(B))                  # d = -grad
        active = np.ones(B, dtype=bool)
        success = np.ones(B, dtype=bool)
        n_success = np.zeros(B, dtype=np.int64)
        beta = np.ones(B)
        kappa, theta, mu = np.zeros(B), np.zeros(B), np.zeros(B)
        fx_out = f_now.copy()
        dfx_prev = self._maxabs_sumabs(self.Gn)[1]                    # sum |g| of the current gradient

        for j in range(nit):
            if not active.any():
                break
            self._rows(active)
            S = success & active
            if S.any():
                # first / second directional derivatives along d  (optim_scg.py:137-170)
                dg, _, dd = self._dot(self.Dd, self.Gn)
                flip = S & (dg >= 0.0)
                if flip.any():
                    self._direction(np.where(flip, 2, 0), np.zeros(B))
                    dg2, _, dd2 = self._dot(self.Dd, self.Gn)
                    dg, dd = np.where(flip, dg2, dg), np.where(flip, dd2, dd)
                mu = np.where(S, dg, mu)
                kappa = np.where(S, dd, kappa)
                tiny = S & (kappa < eps)
                if tiny.any():                                        # optim_scg.py:148-156
                    st["MaxIt"][tiny] = j + 1
                    fx_out[tiny] = f_now[tiny]
                    active &= ~tiny
                    S &= ~tiny
                sigma = np.where(S, self.SIGMA0 / np.sqrt(np.where(kappa > 0, kappa, 1.0)), 0.0)
                self._axpy(sigma, self.Dd, self.X, self.XT)           # x_plus = x + sigma d
                self._eval(self.XT, self.GT, S)                       # df(x_plus, eval_fun=True)
                st["evaluations"] += 1
                st["f_eval"][S] += 1
                st["df_eval"][S] += 1
                dgp, _, _ = self._dot(self.Dd, self.GT)
                theta = np.where(S, (dgp - mu) / np.where(S, sigma, 1.0), theta)
            # effective curvature and step length  (optim_scg.py:173-186)
            delta = theta + beta * kappa
            neg = active & (delta <= 0.0)
            delta = np.where(neg, beta * kappa, delta)
            beta = np.where(neg, beta - theta / np.where(kappa != 0, kappa, 1.0), beta)
            alpha = np.where(active, -(mu / np.where(delta != 0, delta, 1.0)), 0.0)
            self._axpy(alpha, self.Dd, self.X, self.XT)               # x_new = x + alpha d
            f_new = self._eval(self.XT, self.GT, active).copy()       # f(x_new); its gradient is kept
            st["evaluations"] += 1
            st["f_eval"][active] += 1
            # comparison ratio  (optim_scg.py:192-204)
            with np.errstate(divide="ignore", invalid="ignore"):
                Delta = 2.0 * (f_new - f_old) / (alpha * mu)
            succ = active & (Delta >= 0.0)
            fail = active & ~succ
            success = np.where(active, succ, success)
            n_success = n_success + succ
            self._copy_where(succ, self.XT, self.X)
            f_now = np.where(succ, f_new, np.where(fail, f_old, f_now))
            # statistics: the reference records sum|g| of the gradient at the PREVIOUS accepted
            # point on success and of grad_old on failure (optim_scg.py:197-209)
            sum_go = self._maxabs_sumabs(self.Go)[1] if fail.any() else dfx_prev
            st["fx"][j] = np.where(active, f_now, st["fx"][j - 1] if j else f_now)
            st["beta"][j] = beta
            st["dfx"][j] = np.where(succ, dfx_prev, sum_go)
            if self.display and j % 10 == 0:
                print(f" {j}: mean fx={np.mean(f_now):.3f}\tactive={int(active.sum())}")
            # termination and the move to the new point  (optim_scg.py:217-247)
            if succ.any():
                maxd, _ = self._maxabs_sumabs(self.Dd)
                done = succ & (np.abs(alpha) * maxd <= self.x_tol) & (np.abs(f_new - f_old) <= self.f_tol)
                st["MaxIt"][done] = j + 1
                fx_out[done] = f_new[done]
                active &= ~done
                move = succ & ~done
                f_old = np.where(move, f_new, f_old)
                self._copy_where(move, self.Gn, self.Go)              # grad_old = grad_new
                self._copy_where(move, self.GT, self.Gn)              # grad_new = df(x) (x == x_new)
                st["f_eval"][move] += 1                               # the reference re-evaluates f(x), df(x)
                st["df_eval"][move] += 1
                gg_all, ggo, _ = self._dot(self.Gn, self.Gn, self.Go)
                zero = move & np.isclose(gg_all, 0.0)
                st["MaxIt"][zero] = j + 1
                fx_out[zero] = f_now[zero]
                active &= ~zero
                dfx_prev = np.where(move, self._maxabs_sumabs(self.Gn)[1], dfx_prev)
            else:
                move = np.zeros(B, dtype=bool)
                gg_all = ggo = np.zeros(B)
            # scale update  (optim_scg.py:250-257)
            beta = np.where(active & (Delta < 0.25), np.minimum(4.0 * beta, self.BETA_MAX), beta)
            beta = np.where(active & (Delta > 0.75), np.maximum(0.5 * beta, self.BETA_MIN), beta)
            # search direction: Polak-Ribiere, restart after n successes  (optim_scg.py:262-274)
            restart = active & (n_success == self.n)
            n_success = np.where(restart, 0, n_success)
            pr = active & ~restart & move
            with np.errstate(divide="ignore", invalid="ignore"):
                gamma = np.where(pr, np.maximum((ggo - gg_all) / mu, 0.0), 0.0)
            mode = np.where(restart, 2, np.where(pr, 1, 0))
            if mode.any():
                self._direction(mode, gamma)
        fx_out = np.where(active, f_old, fx_out)                      # max_it reached: optim_scg.py:281
        return self.X, fx_out
