"""numpy restatement of the reference's AGQ hot path (TEST INFRASTRUCTURE ONLY).

Functions and the reference code they follow (paths relative to /root/reference):

  gauher          R/Functions.R:306-341
  cd_vec          R/Functions.R:249-262
  dmvnorm         R/Functions.R:264-293
  chol_transf     R/Functions.R:140-158
  deriv_D         R/Functions.R:160-172
  jacobian2       R/Functions.R:174-202
  vmmin           R's optim(method = "BFGS") (src/appl/optim.c, not in the
                  reference tree); call sites
                  R/Functions.R:80 and R/mixed_fit.R:353-355
  find_modes      R/Functions.R:1-103   (BFGS + cd_vec Hessian, as the reference)
  find_modes_newton   same objective/gradient; damped Newton (the engine's
                  algorithm) with the reference's cd_vec Hessian at the mode
  GHfun           R/Functions.R:105-138
  logLik_mixed    R/Fit_Funs.R:1-42
  score_mixed     R/Fit_Funs.R:44-293

The N x K matrices (Ztb, eta_y, log_Lik, p_by, ...) are materialised exactly as
the reference does, so this file also serves as the "port" CPU baseline.
"""
from dataclasses import dataclass, field
from typing import Optional
import numpy as np
from scipy.special import logsumexp

DBL_EPS = np.finfo(float).eps


# ----------------------------------------------------------------------------
# data container (what mixed_fit() receives; R/mixed_fit.R:1-25)
# ----------------------------------------------------------------------------
@dataclass
class Data:
    y: np.ndarray                    # (N,) or (N, 2)
    X: np.ndarray                    # (N, p)
    Z: np.ndarray                    # (N, q_y)
    id: np.ndarray                   # (N,) 0-based, contiguous, ascending (R/mixed_model.R:50,104)
    X_zi: Optional[np.ndarray] = None
    Z_zi: Optional[np.ndarray] = None
    offset: Optional[np.ndarray] = None
    offset_zi: Optional[np.ndarray] = None
    weights: Optional[np.ndarray] = None   # per cluster (n,)
    row_ptr: np.ndarray = field(init=False)

    def __post_init__(self):
        self.y = np.asarray(self.y, float)
        self.X = np.asarray(self.X, float)
        self.Z = np.asarray(self.Z, float)
        if self.Z.ndim == 1:
            self.Z = self.Z[:, None]
        self.id = np.asarray(self.id)
        N = self.X.shape[0]
        change = np.flatnonzero(np.diff(self.id)) + 1
        self.row_ptr = np.concatenate(([0], change, [N])).astype(np.int64)
        assert len(np.unique(self.id)) == len(self.row_ptr) - 1, "id must be contiguous"

    @property
    def n(self):
        return len(self.row_ptr) - 1

    @property
    def N(self):
        return self.X.shape[0]

    @property
    def q_y(self):
        return self.Z.shape[1]

    @property
    def q_zi(self):
        return 0 if self.Z_zi is None else self.Z_zi.shape[1]

    @property
    def nRE(self):
        return self.q_y + self.q_zi

    def y1(self):
        return self.y[:, 0] if self.y.ndim == 2 else self.y


def family_flags(family):
    """canonical / user_defined flags, R/mixed_fit.R:34-38."""
    canonical = (family.family == "binomial" and family.link == "logit") or \
                (family.family == "poisson" and family.link == "log")
    user_defined = family.family not in ("binomial", "poisson")
    return canonical, user_defined


def binomial_N(data, family):
    """N <- if (NCOL(y) == 2 && fams_N) y[, 1] + y[, 2]   (R/mixed_fit.R:8-9)"""
    if data.y.ndim == 2 and family.family in ("binomial", "beta binomial"):
        return data.y[:, 0] + data.y[:, 1]
    return None


# ----------------------------------------------------------------------------
# small utilities
# ----------------------------------------------------------------------------
def gauher(n):
    """R/Functions.R:306-341 (Numerical Recipes gauher; nodes descending)."""
    m = int((n + 1) // 2)
    x = -np.ones(n)
    w = -np.ones(n)
    z = 0.0
    for i in range(1, m + 1):
        if i == 1:
            z = np.sqrt(2 * n + 1) - 1.85575 * (2 * n + 1) ** (-0.16667)
        elif i == 2:
            z = z - 1.14 * n ** 0.426 / z
        elif i == 3:
            z = 1.86 * z - 0.86 * x[0]
        elif i == 4:
            z = 1.91 * z - 0.91 * x[1]
        else:
            z = 2 * z - x[i - 3]
        pp = 0.0
        for _its in range(10):
            p1 = 0.751125544464943
            p2 = 0.0
            for j in range(1, n + 1):
                p3 = p2
                p2 = p1
                p1 = z * np.sqrt(2.0 / j) * p2 - np.sqrt((j - 1.0) / j) * p3
            pp = np.sqrt(2.0 * n) * p2
            z1 = z
            z = z1 - p1 / pp
            if abs(z - z1) <= 3e-14:
                break
        x[i - 1] = z
        x[n - i] = -z
        w[i - 1] = 2.0 / (pp * pp)
        w[n - i] = w[i - 1]
    return x, w


def expand_grid(v, q):
    """as.matrix(expand.grid(rep(list(v), q))): first column varies fastest."""
    k = len(v)
    K = k ** q
    out = np.empty((K, q))
    idx = np.arange(K)
    for d in range(q):
        out[:, d] = v[(idx // (k ** d)) % k]
    return out


def cd_vec(x, f, eps=0.001):
    """R/Functions.R:249-262."""
    x = np.asarray(x, float)
    n = len(x)
    res = np.zeros((n, n))
    ex = np.maximum(np.abs(x), 1.0)
    for i in range(n):
        x1 = x.copy()
        x2 = x.copy()
        x1[i] = x[i] + eps * ex[i]
        x2[i] = x[i] - eps * ex[i]
        diff_f = f(x1) - f(x2)
        diff_x = x1[i] - x2[i]
        res[:, i] = diff_f / diff_x
    return 0.5 * (res + res.T)


def dnorm_log(x, mean, sd):
    z = (x - mean) / sd
    return -0.5 * z * z - np.log(sd) - 0.5 * np.log(2 * np.pi)


def dmvnorm_log(x, mu, Sigma):
    """R/Functions.R:264-293 with log = TRUE."""
    x = np.atleast_2d(np.asarray(x, float))
    mu = np.asarray(mu, float)
    Sigma = np.asarray(Sigma, float)
    p = len(mu)
    if p == 1:
        return dnorm_log(x, mu, np.sqrt(Sigma)).reshape(-1)
    t1 = len(mu) == Sigma.size
    low = Sigma[np.tril_indices(p, -1)] if Sigma.ndim == 2 else np.array([])
    t2 = bool(np.all(np.abs(low) < np.sqrt(DBL_EPS)))
    if t1 or t2:
        sig = Sigma if t1 else np.diag(Sigma)
        return dnorm_log(x, mu[None, :], np.sqrt(sig)[None, :]).sum(axis=1)
    ev, evec = np.linalg.eigh(Sigma)
    ev, evec = ev[::-1], evec[:, ::-1]      # eigen(): decreasing order
    if not np.all(ev >= -1e-06 * abs(ev[0])):
        raise ValueError("'Sigma' is not positive definite")
    ss = x - mu[None, :]
    inv_Sigma = evec @ (evec.T / ev[:, None])
    quad = 0.5 * np.sum((ss @ inv_Sigma) * ss, axis=1)
    fact = -0.5 * (p * np.log(2 * np.pi) + np.sum(np.log(ev)))
    return fact - quad


def _upper_tri_colmajor(k):
    """index pairs of M[upper.tri(M, TRUE)] in R's column-major order."""
    return [(r, c) for c in range(k) for r in range(c + 1)]


def _lower_tri_colmajor(k):
    return [(r, c) for c in range(k) for r in range(c, k)]


def chol_transf(x):
    """R/Functions.R:140-158.  Matrix -> unconstrained vector; vector -> (D, L)."""
    x = np.asarray(x, float)
    if np.any(~np.isfinite(x)):
        raise ValueError("NA or infinite values in 'x'.")
    if x.ndim == 2:
        k = x.shape[0]
        U = np.linalg.cholesky(x).T.copy()      # chol(): upper, U'U = x
        U[np.arange(k), np.arange(k)] = np.log(np.diag(U))
        return np.array([U[r, c] for r, c in _upper_tri_colmajor(k)])
    nx = len(x)
    k = int(round((-1 + np.sqrt(1 + 8 * nx)) / 2))
    mat = np.zeros((k, k))
    for v, (r, c) in zip(x, _upper_tri_colmajor(k)):
        mat[r, c] = v
    mat[np.arange(k), np.arange(k)] = np.exp(np.diag(mat))
    res = mat.T @ mat                            # crossprod(mat)
    tm = mat.T
    L = np.array([tm[r, c] for r, c in _lower_tri_colmajor(k)])
    return res, L


def deriv_D(D):
    """R/Functions.R:160-172."""
    ncz = D.shape[0]
    out = []
    for r, c in _lower_tri_colmajor(ncz):
        mat = np.zeros((ncz, ncz))
        mat[c, r] = mat[r, c] = 1.0
        out.append(mat)
    return out


def jacobian2(L, ncz):
    """R/Functions.R:174-202, translated with 1-based index emulation."""
    ind = [(r + 1, c + 1) for r, c in _lower_tri_colmajor(ncz)]   # (row, col), 1-based
    nind = len(ind)
    ids = list(range(1, nind + 1))
    rind = [i for i in ids if ind[i - 1][0] == ind[i - 1][1]]
    lind = []
    for i in range(1, len(rind) + 1):
        m = ncz - i + 1
        tt = np.zeros((m, m))
        seq = list(range(rind[i - 1], nind + 1))
        for v, (r, c) in zip(seq, _lower_tri_colmajor(m)):
            tt[r, c] = v
        tt = tt + tt.T
        tt[np.arange(m), np.arange(m)] = np.diag(tt) / 2
        lind.append(tt)
    out = np.zeros((nind, nind))
    L = np.asarray(L, float)
    for g in range(1, ncz + 1):
        gind = [i for i in ids if ind[i - 1][1] == g]
        vals = L[[i - 1 for i in gind]]
        for j in gind:
            k = gind.index(j)
            rows = lind[g - 1][k, :].astype(int)
            v = vals[0] * vals if j in rind else vals
            for rr, vv in zip(rows, v):
                out[rr - 1, j - 1] = vv
    for r in rind:
        out[r - 1, :] = 2 * out[r - 1, :]
    col_ind = np.zeros((ncz, ncz))
    for v, (r, c) in zip(range(1, len(L) + 1), _lower_tri_colmajor(ncz)):
        col_ind[r, c] = v
    col_ind = col_ind.T
    cols = [int(col_ind[r, c]) - 1 for r, c in _upper_tri_colmajor(ncz)]
    return out[:, cols]


# ----------------------------------------------------------------------------
# optim(method = "BFGS")
# ----------------------------------------------------------------------------
def vmmin(par, fn, gr, maxit=100, reltol=np.sqrt(DBL_EPS), abstol=-np.inf, parscale=None):
    """R's vmmin.  Returns dict(par, value, counts, convergence)."""
    stepredn, acctol, reltest = 0.2, 1e-4, 10.0
    par = np.asarray(par, float)
    n = len(par)
    ps = np.ones(n) if parscale is None else np.asarray(parscale, float)
    fminfn = lambda b: float(fn(b * ps))
    fmingr = lambda b: np.asarray(gr(b * ps), float) * ps
    b = par / ps
    if maxit <= 0:
        return dict(par=par, value=fminfn(b), counts=(0, 0), convergence=0)
    f = fminfn(b)
    if not np.isfinite(f):
        raise FloatingPointError("initial value in 'vmmin' is not finite")
    Fmin = f
    funcount = gradcount = 1
    g = fmingr(b)
    it = 1
    ilast = gradcount
    B = np.eye(n)
    count = 0
    while True:
        if ilast == gradcount:
            B = np.eye(n)
        X = b.copy()
        c = g.copy()
        t = -B @ g
        gradproj = float(t @ g)
        if gradproj < 0.0:
            steplength = 1.0
            accpoint = False
            while True:
                b = X + steplength * t
                count = int(np.sum((reltest + X) == (reltest + b)))
                if count < n:
                    f = fminfn(b)
                    funcount += 1
                    accpoint = np.isfinite(f) and (f <= Fmin + gradproj * steplength * acctol)
                    if not accpoint:
                        steplength *= stepredn
                if count == n or accpoint:
                    break
            enough = (f > abstol) and abs(f - Fmin) > reltol * (abs(Fmin) + reltol)
            if not enough:
                count = n
                Fmin = f
            if count < n:
                Fmin = f
                g = fmingr(b)
                gradcount += 1
                it += 1
                t = steplength * t
                c = g - c
                D1 = float(t @ c)
                if D1 > 0:
                    Xv = B @ c
                    D2 = 1.0 + float(Xv @ c) / D1
                    B = B + (D2 * np.outer(t, t) - np.outer(Xv, t) - np.outer(t, Xv)) / D1
                else:
                    ilast = gradcount
            else:
                if ilast < gradcount:
                    count = 0
                    ilast = gradcount
        else:
            count = 0
            if ilast == gradcount:
                count = n
            else:
                ilast = gradcount
        if it >= maxit:
            break
        if gradcount - ilast > 2 * n:
            ilast = gradcount
        if not (count != n or ilast != gradcount):
            break
    return dict(par=b * ps, value=Fmin, counts=(funcount, gradcount),
                convergence=0 if it < maxit else 1)


# ----------------------------------------------------------------------------
# find_modes / GHfun
# ----------------------------------------------------------------------------
def _cluster_views(data, i):
    lo, hi = data.row_ptr[i], data.row_ptr[i + 1]
    sl = slice(lo, hi)
    g = lambda a: None if a is None else a[sl]
    return g(data.y), g(data.X), g(data.Z), g(data.offset), g(data.X_zi), g(data.Z_zi), g(data.offset_zi)


def make_log_post(data, family, i, betas, invD, phis, gammas):
    """log_post_b and score_log_post_b for cluster i (R/Functions.R:5-65)."""
    canonical, user_defined = family_flags(family)
    y_i, X_i, Z_i, off_i, Xzi_i, Zzi_i, offzi_i = _cluster_views(data, i)
    ncz = Z_i.shape[1]
    N_all = binomial_N(data, family)
    N_i = None if N_all is None else N_all[data.row_ptr[i]:data.row_ptr[i + 1]]
    y1_i = y_i[:, 0] if y_i.ndim == 2 else y_i
    Zty_i = Z_i.T @ y1_i

    def etas(b_i):
        eta_y = X_i @ betas + Z_i @ b_i[:ncz]
        if off_i is not None:
            eta_y = eta_y + off_i
        eta_zi = None
        if Xzi_i is not None:
            eta_zi = Xzi_i @ gammas
        if Zzi_i is not None:
            eta_zi = eta_zi + Zzi_i @ b_i[ncz:]
        if offzi_i is not None:
            eta_zi = eta_zi + offzi_i
        return eta_y, eta_zi

    def log_post_b(b_i):
        eta_y, eta_zi = etas(b_i)
        ld, _ = family.log_dens(y_i, eta_y, phis, eta_zi)
        return -np.nansum(ld) + 0.5 * float(b_i @ invD @ b_i)

    def score_log_post_b(b_i):
        eta_y, eta_zi = etas(b_i)
        mu_y = family.mu_fun(eta_y)
        if user_defined:
            s = family.score_eta_fun(y_i, mu_y, phis, eta_zi)
            out = -(Z_i.T @ np.reshape(s, -1))
            if Zzi_i is not None:
                sz = family.score_eta_zi_fun(y_i, mu_y, phis, eta_zi)
                out = np.concatenate([out, -(Zzi_i.T @ np.reshape(sz, -1))])
        else:
            if canonical:
                out = -Zty_i + (Z_i.T @ (N_i * mu_y) if N_i is not None else Z_i.T @ mu_y)
            else:
                var = family.variance(mu_y)
                deriv = family.mu_eta(eta_y)
                if N_i is not None:
                    out = -(Z_i.T @ ((y_i[:, 0] - N_i * mu_y) * deriv / var))
                else:
                    out = -(Z_i.T @ ((y_i - mu_y) * deriv / var))
        return out + invD @ b_i

    return log_post_b, score_log_post_b


def find_modes(b, data, family, betas, invD, phis, gammas):
    """R/Functions.R:1-103: per-cluster optim(BFGS) + cd_vec Hessian."""
    n = data.n
    post_modes = np.array(b, float, copy=True)
    post_hessians = []
    for i in range(n):
        fn, gr = make_log_post(data, family, i, betas, invD, phis, gammas)
        opt = vmmin(post_modes[i], fn, gr)
        post_modes[i] = opt["par"]
        post_hessians.append(cd_vec(post_modes[i], gr))
    return post_modes, post_hessians


def find_modes_newton(b, data, family, betas, invD, phis, gammas, tol=1e-10, maxit=100):
    """The engine's mode search restated on the CPU: damped Newton on the
    reference's objective (R/Functions.R:5-20) with its analytic gradient
    (:21-65); the Newton matrix is the reference's own cd_vec Hessian of that
    gradient (:90, :249-262), Levenberg-shifted if not PD, Armijo backtracking
    (c = 1e-4, halving, plus a rounding slack of 1e-14 (|f| + 1) so that full steps
    inside the quadratic region are not rejected by noise in f).  Stops when the max-norm of the Newton step is
    <= tol.  Returns modes, Hessians (cd_vec at the mode) and a convergence flag."""
    n = data.n
    post_modes = np.array(b, float, copy=True)
    post_hessians = []
    conv = np.zeros(n, bool)
    for i in range(n):
        fn, gr = make_log_post(data, family, i, betas, invD, phis, gammas)
        x = post_modes[i].copy()
        fx = fn(x)
        for _ in range(maxit):
            g = gr(x)
            H = cd_vec(x, gr)
            lam = 0.0
            while True:
                try:
                    Lc = np.linalg.cholesky(H + lam * np.eye(len(x)))
                    break
                except np.linalg.LinAlgError:
                    lam = max(10 * lam, 1e-3 * max(1.0, np.max(np.abs(np.diag(H)))))
            d = -np.linalg.solve(Lc.T, np.linalg.solve(Lc, g))
            step = 1.0
            gd = float(g @ d)
            ok = False
            for _h in range(60):
                xn = x + step * d
                fnew = fn(xn)
                if np.isfinite(fnew) and fnew <= fx + 1e-4 * step * gd + 1e-14 * (abs(fx) + 1.0):
                    ok = True
                    break
                step *= 0.5
            if not ok:
                conv[i] = np.max(np.abs(d)) <= tol      # a step below tol that rounding noise refuses
                break
            x, fx = xn, fnew
            if np.max(np.abs(d)) <= tol:
                conv[i] = True
                break
        post_modes[i] = x
        post_hessians.append(cd_vec(x, gr))
    return post_modes, post_hessians, conv


@dataclass
class GH:
    """The list GHfun returns (R/Functions.R:135-137)."""
    b: np.ndarray            # (n*K, q)   cluster-major, node-minor
    b2: np.ndarray           # (n*K, q*q)
    Ztb: np.ndarray          # (N, K)
    Z_zitb: Optional[np.ndarray]
    wGH: np.ndarray          # (K,)
    log_dets: np.ndarray     # (n,)
    post_modes: np.ndarray   # (n, q)
    post_vars: list
    chol_hessians: list = None   # kept for tests (not in the reference's list)


def GHfun_from_modes(data, modes, hessians, k, q):
    """R/Functions.R:114-137 given modes and Hessians."""
    x, w = gauher(k)
    chol_hessians = [np.linalg.cholesky(H).T for H in hessians]    # chol(): upper
    b = expand_grid(x, q)
    n = modes.shape[0]
    K = b.shape[0]
    b_new = np.empty((n, K, q))
    log_dets = np.empty(n)
    for i in range(n):
        R = chol_hessians[i]
        b_new[i] = (np.sqrt(2) * np.linalg.solve(R, b.T) + modes[i][:, None]).T
        log_dets[i] = -np.sum(np.log(np.abs(np.diag(R))))
    wg = expand_grid(w, q)
    wGH = 2 ** (q / 2) * np.prod(wg, axis=1) * np.exp(np.sum(b * b, axis=1))
    b2 = (b_new[:, :, :, None] * b_new[:, :, None, :]).reshape(n * K, q * q)
    ncz = data.q_y
    Ztb = np.empty((data.N, K))
    Z_zitb = None if data.Z_zi is None else np.empty((data.N, K))
    for i in range(n):
        lo, hi = data.row_ptr[i], data.row_ptr[i + 1]
        Ztb[lo:hi] = data.Z[lo:hi] @ b_new[i][:, :ncz].T
        if data.Z_zi is not None:
            Z_zitb[lo:hi] = data.Z_zi[lo:hi] @ b_new[i][:, ncz:].T
    return GH(b=b_new.reshape(n * K, q), b2=b2, Ztb=Ztb, Z_zitb=Z_zitb, wGH=wGH,
              log_dets=log_dets, post_modes=modes,
              post_vars=[np.linalg.inv(H) for H in hessians], chol_hessians=chol_hessians)


def GHfun(b, data, family, betas, inv_D, phis, gammas, k, q, mode_finder="bfgs"):
    """R/Functions.R:105-138.  mode_finder = "bfgs" (reference) or "newton" (engine)."""
    if mode_finder == "bfgs":
        modes, hess = find_modes(b, data, family, betas, inv_D, phis, gammas)
    else:
        modes, hess, _ = find_modes_newton(b, data, family, betas, inv_D, phis, gammas)
    return GHfun_from_modes(data, modes, hess, k, q)


# ----------------------------------------------------------------------------
# logLik_mixed / score_mixed
# ----------------------------------------------------------------------------
def split_thetas(thetas, p, nRE, n_phis, p_zi, diag_D):
    """relist(thetas, skeleton = list_thetas): c(betas, D, phis, gammas) (R/mixed_fit.R:240-247)."""
    thetas = np.asarray(thetas, float)
    nD = nRE if diag_D else nRE * (nRE + 1) // 2
    o = 0
    betas = thetas[o:o + p]; o += p
    Dp = thetas[o:o + nD]; o += nD
    phis = thetas[o:o + n_phis] if n_phis else None; o += n_phis
    gammas = thetas[o:o + p_zi] if p_zi else None
    return betas, Dp, phis, gammas


def _theta_parts(thetas, data, family, diag_D):
    p = data.X.shape[1]
    p_zi = 0 if data.X_zi is None else data.X_zi.shape[1]
    betas, Dp, phis, gammas = split_thetas(thetas, p, data.nRE, family.n_phis, p_zi, diag_D)
    if diag_D:
        D = np.diag(np.exp(Dp))
        L = None
    else:
        D, L = chol_transf(Dp)
    return betas, D, L, phis, gammas


def _common(thetas, data, family, gh, diag_D):
    betas, D, L, phis, gammas = _theta_parts(thetas, data, family, diag_D)
    nRE = D.shape[0]
    eta_y = (data.X @ betas)[:, None] + gh.Ztb
    if data.offset is not None:
        eta_y = eta_y + data.offset[:, None]
    eta_zi = None
    if data.X_zi is not None:
        eta_zi = data.X_zi @ gammas
    if data.Z_zi is not None:
        eta_zi = eta_zi[:, None] + gh.Z_zitb
    if data.offset_zi is not None:
        eta_zi = eta_zi + (data.offset_zi[:, None] if eta_zi.ndim == 2 else data.offset_zi)
    log_Lik, mu_y = family.log_dens(data.y, eta_y, phis, eta_zi)
    log_p_yb = np.add.reduceat(log_Lik, data.row_ptr[:-1], axis=0)      # rowsum(., id, reorder = FALSE)
    n, K = log_p_yb.shape
    log_p_b = dmvnorm_log(gh.b, np.zeros(nRE), D if nRE > 1 else D.reshape(1, 1)[0]).reshape(n, K)
    log_wGH = np.log(gh.wGH)[None, :]
    return dict(betas=betas, D=D, L=L, phis=phis, gammas=gammas, nRE=nRE, eta_y=eta_y,
                eta_zi=eta_zi, log_Lik=log_Lik, mu_y=mu_y, log_p_yb=log_p_yb, log_p_b=log_p_b,
                log_wGH=log_wGH, n=n, K=K)


class Penalty:
    """Student's t penalty on the fixed effects: R/mixed_model.R:196-208 (pen_mu = 0, pen_sigma = 1,
    pen_df = 3 for penalized = TRUE), R/mixed_fit.R:64-69 (pen_invSigma = diag(1 / pen_sigma^2)),
    dmvt(log = TRUE, prop = TRUE) of R/Functions.R:424-462 and the score term of
    R/Fit_Funs.R:169-173 / :360-364 (which uses betas, not betas - pen_mu: the reference's form)."""

    def __init__(self, ncx, pen_mu=0.0, pen_sigma=1.0, pen_df=3.0):
        self.mu = np.resize(np.atleast_1d(np.asarray(pen_mu, float)), ncx)          # rep(., length.out = ncx)
        self.inv = np.resize(1.0 / np.atleast_1d(np.asarray(pen_sigma, float)) ** 2, ncx)
        self.df = float(pen_df)

    def log_dens(self, betas):
        ss = np.asarray(betas, float) - self.mu
        quad = np.sum(ss * self.inv * ss) / self.df
        return -0.5 * (self.df + len(ss)) * np.log(1.0 + quad)

    def score(self, betas):
        betas = np.asarray(betas, float)
        v = betas * self.inv / self.df
        return v * (self.df + len(betas)) / (1.0 + betas @ v)


def rowsum(a, data):
    return np.add.reduceat(a, data.row_ptr[:-1], axis=0)


def logLik_mixed(thetas, data, family, gh, diag_D=False, i_contributions=False, penalty=None):
    """R/Fit_Funs.R:1-42.  Returns the NEGATIVE log-likelihood (minus the log penalty, :39-40)."""
    c = _common(thetas, data, family, gh, diag_D)
    log_p_y = logsumexp(c["log_p_yb"] + c["log_p_b"] + c["log_wGH"], axis=1) + gh.log_dets
    w = data.weights
    out = -(log_p_y if w is None else w * log_p_y) if i_contributions else \
        -np.nansum(log_p_y if w is None else w * log_p_y)
    if penalty is not None:
        out = out - penalty.log_dens(c["betas"])
    return out


def score_mixed(thetas, data, family, gh, diag_D=False, i_contributions=False, return_parts=False,
                penalty=None):
    """R/Fit_Funs.R:44-293.  Gradient of the NEGATIVE log-likelihood (plus the penalty term, :169-173),
    c(score.betas, score.D, score.phis, score.gammas)."""
    canonical, user_defined = family_flags(family)
    c = _common(thetas, data, family, gh, diag_D)
    D, nRE, n, K = c["D"], c["nRE"], c["n"], c["K"]
    phis, eta_zi, eta_y = c["phis"], c["eta_zi"], c["eta_y"]
    w = data.weights
    idx = np.repeat(np.arange(n), np.diff(data.row_ptr))
    log_p_yb_b = c["log_p_yb"] + c["log_p_b"]
    log_p_y = logsumexp(log_p_yb_b + c["log_wGH"], axis=1)
    p_by = np.exp(log_p_yb_b - log_p_y[:, None])
    wGH = gh.wGH
    bb = gh.b.reshape(n, K, nRE)
    b2 = gh.b2.reshape(n, K, nRE * nRE)
    post_b = np.einsum("ik,ikd,k->id", p_by, bb, wGH)
    post_b2 = np.einsum("ik,ikd,k->id", p_by, b2, wGH)
    if w is not None:
        post_b = w[:, None] * post_b
        post_b2 = w[:, None] * post_b2
    post_vb = post_b2 - (post_b[:, :, None] * post_b[:, None, :]).reshape(n, nRE * nRE)
    mu_y = c["mu_y"]
    X = data.X
    ncx = X.shape[1]
    y = data.y
    N = binomial_N(data, family)

    def col_scores(Xm, z):
        ncol = Xm.shape[1]
        sc = np.zeros((data.N, ncol)) if i_contributions else np.zeros(ncol)
        for l in range(ncol):
            if i_contributions:
                sc[:, l] = (Xm[:, [l]] * z * p_by[idx, :]) @ wGH
            else:
                cc = rowsum(Xm[:, [l]] * z, data)
                if w is not None:
                    cc = w[:, None] * cc
                sc[l] = np.nansum((cc * p_by) @ wGH)
        return sc

    if user_defined:
        z = family.score_eta_fun(y, mu_y, phis, eta_zi)
        score_betas = -col_scores(X, z)
    else:
        if canonical:
            mu_n = mu_y * N[:, None] if N is not None else mu_y
            sc = col_scores(X, mu_n)
            y1 = data.y1()
            if i_contributions:
                score_betas = -(X * y1[:, None]) + sc
            else:
                Xty = X.T @ y1 if w is None else X.T @ (y1 * w[idx])
                score_betas = -Xty + sc
        else:
            var = family.variance(mu_y)
            deriv = family.mu_eta(eta_y)
            if N is not None:
                z = (y[:, [0]] - N[:, None] * mu_y) * deriv / var
            else:
                z = (y[:, None] - mu_y) * deriv / var
            score_betas = -col_scores(X, z)

    if penalty is not None:                   # R/Fit_Funs.R:169-173
        score_betas = score_betas + penalty.score(c["betas"])

    score_phis = None
    if phis is not None:
        z = family.score_phis_fun(y, mu_y, phis, eta_zi)
        if i_contributions:
            score_phis = -((z * p_by[idx, :]) @ wGH)
        else:
            cc = (rowsum(z, data) * p_by) @ wGH
            if w is not None:
                cc = w * cc
            score_phis = np.array([-np.nansum(cc)])

    score_gammas = None
    if data.X_zi is not None:
        z = family.score_eta_zi_fun(y, mu_y, phis, eta_zi)
        score_gammas = -col_scores(data.X_zi, z)

    if diag_D:
        Dd = np.diag(D)
        svD = 1.0 / Dd
        svD2 = svD ** 2
        if i_contributions:
            score_D = None
        else:
            cS = np.nansum(post_vb, axis=0).reshape(nRE, nRE)
            score_D = Dd * 0.5 * (n * svD - np.diag(cS) * svD2 - np.nansum(post_b ** 2, axis=0) * svD2)
    else:
        svD = np.linalg.inv(D)
        dD = deriv_D(D)
        ndD = len(dD)
        D1 = np.array([np.sum(svD * x) for x in dD])
        # c() of a matrix is column-major; svD x svD is symmetric so the order is immaterial
        D2 = np.array([(svD @ x @ svD).reshape(-1, order="F") for x in dD])
        J = jacobian2(c["L"], nRE)
        if i_contributions:
            rr = np.zeros((n, ndD))
            for j in range(n):
                out = np.zeros(ndD)
                for i in range(ndD):
                    Dmat = D2[i].reshape(nRE, nRE, order="F")
                    out[i] = np.nansum(D2[i] * post_vb[j]) + np.nansum((post_b[[j]] @ Dmat) * post_b[[j]])
                rr[j] = 0.5 * (D1 - out) @ J
            score_D = rr
        else:
            cS = np.nansum(post_vb, axis=0)
            out = np.zeros(ndD)
            for i in range(ndD):
                Dmat = D2[i].reshape(nRE, nRE, order="F")
                out[i] = np.nansum(D2[i] * cS) + np.nansum((post_b @ Dmat) * post_b)
            sw = n if w is None else np.sum(w)
            score_D = 0.5 * (sw * D1 - out) @ J

    if i_contributions or return_parts:
        return dict(score_betas=score_betas, score_D=score_D, score_phis=score_phis,
                    score_gammas=score_gammas, p_by=p_by, post_b=post_b, post_b2=post_b2)
    parts = [score_betas, score_D]
    if score_phis is not None:
        parts.append(score_phis)
    if score_gammas is not None:
        parts.append(score_gammas)
    return np.concatenate(parts)


def suff_stats(thetas, data, family, gh, diag_D=False):
    """What the engine's C ABI returns for one evaluation (include/glmma.h,
    glmma_result), computed from the reference-shaped quantities above:
    loglik = sum_i w_i l_i, the positive score sums g_beta/g_phi/g_gamma
    (= minus the reference's score.* before the D Jacobian) and
    S = sum_i w_i E[b b' | y_i]  (SURVEY.md section 8a row a12)."""
    c = _common(thetas, data, family, gh, diag_D)
    canonical, user_defined = family_flags(family)
    n, K, nRE = c["n"], c["K"], c["nRE"]
    w = np.ones(n) if data.weights is None else data.weights
    a = c["log_p_yb"] + c["log_p_b"] + c["log_wGH"]
    lse = logsumexp(a, axis=1)
    ll_i = lse + gh.log_dets
    post = np.exp(a - lse[:, None])                     # pi_ik
    idx = np.repeat(np.arange(n), np.diff(data.row_ptr))
    mu_y = c["mu_y"]
    N = binomial_N(data, family)
    if user_defined:
        s_eta = family.score_eta_fun(data.y, mu_y, c["phis"], c["eta_zi"])
    elif canonical:
        y1 = data.y1()[:, None]
        s_eta = y1 - (mu_y * N[:, None] if N is not None else mu_y)
    else:
        var = family.variance(mu_y)
        deriv = family.mu_eta(c["eta_y"])
        y1 = data.y1()[:, None]
        s_eta = (y1 - (N[:, None] * mu_y if N is not None else mu_y)) * deriv / var
    ok = np.isfinite(ll_i)
    wz = (w * ok)[idx][:, None] * post[idx, :]
    zbar = np.nansum(wz * s_eta, axis=1)
    out = dict(loglik=float(np.sum((w * ll_i)[ok])), ll_i=ll_i, g_beta=data.X.T @ zbar,
               n_nonfinite=int(np.sum(~ok)), sum_w=float(np.sum(w[ok])), zbar=zbar)
    if c["phis"] is not None:
        sp_ = family.score_phis_fun(data.y, mu_y, c["phis"], c["eta_zi"])
        out["g_phi"] = np.array([np.nansum(wz * sp_)])
    if data.X_zi is not None:
        sz = family.score_eta_zi_fun(data.y, mu_y, c["phis"], c["eta_zi"])
        out["g_gamma"] = data.X_zi.T @ np.nansum(wz * sz, axis=1)
    bb = gh.b.reshape(n, K, nRE)
    S_i = np.einsum("ik,ikd,ike->ide", post, bb, bb)
    out["S_i"] = S_i
    out["S"] = np.einsum("i,ide->de", w * ok, np.where(ok[:, None, None], S_i, 0.0))
    return out
