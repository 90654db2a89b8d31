"""
Scaled Conjugate Gradient optimiser with the interface and the statistics of the
reference's src/numerics/optim_scg.py:23-295 (NETLAB's scg, Nabney 2001; Moller 1993):
`SCG(f, df, {"max_it", "x_tol", "f_tol", "display"})(x0) -> (x, fx)` and
`.stats = {MaxIt, fx[], dfx[], beta[], f_eval, df_eval}`.

It is the CALLER of the hot path, kept on the host; the arithmetic is arranged as in
the reference so that identical f/df values give the identical convergence trace
(the 1e-6 trace criterion of BASELINE.json is then a statement about F and grad F).
"""
import numpy as np


class SCG(object):
    SIGMA0 = 1.0e-3
    BETA_MIN, BETA_MAX = 1.0e-15, 1.0e+100

    def __init__(self, f, df, *args):
        opts = args[0] if args else {}
        self.f, self.df = f, df
        self.nit = opts.get("max_it", 150)
        self.x_tol = opts.get("x_tol", 1.0e-6)
        self.f_tol = opts.get("f_tol", 1.0e-8)
        self.display = opts.get("display", False)
        self.stats = {"MaxIt": self.nit, "fx": np.zeros(self.nit), "dfx": np.zeros(self.nit),
                      "f_eval": 0.0, "df_eval": 0.0, "beta": np.zeros(self.nit)}

    @property
    def statistics(self):
        return self.stats

    def __call__(self, x0, *args):
        st = self.stats
        x = x0.flatten()
        n = x.size
        eps = np.finfo(float).eps
        f_now = self.f(x, *args)
        g_new = self.df(x, *args)
        st["f_eval"] += 1
        st["df_eval"] += 1
        f_old, g_old = f_now, g_new.copy()
        d = -g_new
        success, n_success = True, 0
        beta, kappa, theta, mu = 1.0, 0.0, 0.0, 0.0
        for j in range(self.nit):
            if success:
                # first and second directional derivatives along d
                mu = d.T.dot(g_new)
                if mu >= 0.0:
                    d = -g_new
                    mu = d.T.dot(g_new)
                kappa = d.T.dot(d)
                if kappa < eps:
                    st["MaxIt"] = j + 1
                    return x, f_now
                sigma = self.SIGMA0 / np.sqrt(kappa)
                g_plus = self.df(x + (sigma * d), eval_fun=True)
                st["f_eval"] += 1
                st["df_eval"] += 1
                theta = (d.T.dot(g_plus - g_new)) / sigma
            # effective curvature, step length
            delta = theta + (beta * kappa)
            if delta <= 0.0:
                delta = beta * kappa
                beta = beta - (theta / kappa)
            alpha = -(mu / delta)
            x_new = x + (alpha * d)
            f_new = self.f(x_new, *args)
            st["f_eval"] += 1
            # comparison ratio
            delta = 2.0 * (f_new - f_old) / (alpha * mu)
            if delta >= 0.0:
                success = True
                n_success += 1
                x, f_now, g_now = x_new.copy(), f_new, g_new.copy()
            else:
                success = False
                f_now, g_now = f_old, g_old.copy()
            total_grad = np.sum(np.abs(g_now))
            st["fx"][j], st["beta"][j], st["dfx"][j] = f_now, beta, total_grad
            if self.display and (j % 10 == 0):
                print(" {0}: fx={1:.3f}\tsum(gx)={2:.3f}".format(j, f_now, total_grad))
            if success:
                if (np.abs(alpha * d).max() <= self.x_tol) and (np.abs(f_new - f_old) <= self.f_tol):
                    st["MaxIt"] = j + 1
                    return x, f_new
                f_old, g_old = f_new, g_new.copy()
                f_now = self.f(x, *args)
                g_new = self.df(x, *args)
                st["f_eval"] += 1
                st["df_eval"] += 1
                if np.isclose(g_new.T.dot(g_new), 0.0):
                    st["MaxIt"] = j + 1
                    return x, f_now
            # trust-region style update of the scale
            if delta < 0.25:
                beta = np.minimum(4.0 * beta, self.BETA_MAX)
            if delta > 0.75:
                beta = np.maximum(0.5 * beta, self.BETA_MIN)
            # Polak-Ribiere direction, restart after n successes
            if n_success == n:
                d = -g_new
                n_success = 0
            elif success:
                gamma = np.maximum(g_new.T.dot(g_old - g_new) / mu, 0.0)
                d = (gamma * d) - g_new
        print(" SGC: Maximum number of iterations has been reached.")
        return x, f_old
