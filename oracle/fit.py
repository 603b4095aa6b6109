"""TEST ORACLE (CPU, numpy): restatement of the reference's fitting loop, so that full
mixed_model() fits of the CUDA engine can be checked end to end.  Test infrastructure only
(imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs; never by the product).

Follows, with file:line of the reference (drizopoulos/GLMMadaptive v0.9-7):

  glm_fit()           stats::glm.fit as called at R/mixed_model.R:162-194 (IRLS; third-party R
                      code, restated from its documented algorithm: family$initialize starts,
                      working response/weights, |dev - dev_old| / (|dev| + 0.1) < 1e-8, maxit 25)
  initial_values()    R/mixed_model.R:157-194, 251-263
  fd_vec(), cd_vec()  R/Functions.R:219-232, 249-262
  nearPD()            R/Functions.R:343-382
  score_betas/phis/gammas with p_by held fixed   R/Fit_Funs.R:295-427
  mixed_fit()         R/mixed_fit.R:62-349 (EM phase :87-229, quasi-Newton phase :239-312,
                      final GH / logLik / Hessian :313-349)

Parity unpinned by reference tests (the reference has none); pinned on the coefficients the
rendered vignettes publish for `fm1` (docs/articles/GLMMadaptive.html) -- tests/test_fit_oracle.py.
"""
import numpy as np
from scipy.special import logsumexp

from . import agq

DBL_EPS = np.finfo(float).eps

CONTROL_DEFAULTS = dict(iter_EM=30, iter_qN_outer=15, iter_qN=10, iter_qN_incr=10, parscale_betas=0.1,
                        parscale_D=0.01, parscale_phis=0.01, parscale_gammas=0.01, tol1=1e-4, tol2=1e-5,
                        tol3=1e-8, numeric_deriv="fd", nAGQ=None, update_GH_every=10, max_coef_value=12.0,
                        max_phis_value=np.exp(10.0))      # R/mixed_model.R:150-156


# ----------------------------------------------------------------------------------------
# stats::glm.fit (IRLS) for the families mixed_model() uses for starting values
# ----------------------------------------------------------------------------------------
def glm_fit(X, y, family, offset=None, maxit=25, epsilon=1e-8):
    X = np.asarray(X, float)
    y = np.asarray(y, float)
    N = X.shape[0]
    off = np.zeros(N) if offset is None else np.asarray(offset, float)
    pw = np.ones(N)
    if family == "binomial":
        if y.ndim == 2:                                   # cbind(successes, failures)
            ntr = y[:, 0] + y[:, 1]
            yy = np.where(ntr > 0, y[:, 0] / np.where(ntr > 0, ntr, 1.0), 0.0)
            pw = ntr
        else:
            yy = y
        mustart = (pw * yy + 0.5) / (pw + 1.0)
        linkfun = lambda m: np.log(m / (1.0 - m))
        linkinv = lambda e: 1.0 / (1.0 + np.exp(-np.clip(e, -30, 30)))   # make.link("logit") clamp ~ +-36
        mu_eta = lambda e: np.maximum(np.exp(-np.abs(e)) / (1.0 + np.exp(-np.abs(e))) ** 2, DBL_EPS)
        variance = lambda m: m * (1.0 - m)

        def dev(yv, m, w):
            with np.errstate(divide="ignore", invalid="ignore"):
                t1 = np.where(yv > 0, yv * np.log(yv / m), 0.0)
                t2 = np.where(yv < 1, (1 - yv) * np.log((1 - yv) / (1 - m)), 0.0)
            return 2.0 * np.sum(w * (t1 + t2))
    elif family == "poisson":
        yy = y
        mustart = yy + 0.1
        linkfun, linkinv = np.log, lambda e: np.maximum(np.exp(e), DBL_EPS)
        mu_eta = lambda e: np.maximum(np.exp(e), DBL_EPS)
        variance = lambda m: m

        def dev(yv, m, w):
            with np.errstate(divide="ignore", invalid="ignore"):
                t = np.where(yv > 0, yv * np.log(yv / m), 0.0)
            return 2.0 * np.sum(w * (t - (yv - m)))
    elif family == "Gamma":                               # stats::Gamma(): inverse link
        yy = y
        mustart = yy
        linkfun, linkinv = (lambda m: 1.0 / m), (lambda e: 1.0 / e)
        mu_eta = lambda e: -1.0 / e ** 2
        variance = lambda m: m ** 2
        dev = lambda yv, m, w: -2.0 * np.sum(w * (np.log(np.where(yv == 0, 1.0, yv / m)) - (yv - m) / m))
    elif family == "gaussian":
        yy = y
        mustart = yy
        linkfun = linkinv = lambda v: v
        mu_eta = lambda e: np.ones_like(e)
        variance = lambda m: np.ones_like(m)
        dev = lambda yv, m, w: np.sum(w * (yv - m) ** 2)
    else:
        raise ValueError(family)
    eta = linkfun(mustart)
    mu = linkinv(eta)
    devold = dev(yy, mu, pw)
    coef = np.zeros(X.shape[1])
    for _ in range(maxit):
        me = mu_eta(eta)
        z = (eta - off) + (yy - mu) / me
        w = np.sqrt(pw * me ** 2 / variance(mu))
        coef = np.linalg.lstsq(X * w[:, None], z * w, rcond=None)[0]
        eta = X @ coef + off
        mu = linkinv(eta)
        d = dev(yy, mu, pw)
        if abs(d - devold) / (abs(d) + 0.1) < epsilon:
            break
        devold = d
    return coef


def initial_values(data, family, diag_D=False):
    """R/mixed_model.R:157-194 and :251-263."""
    nRE = data.nRE
    known = ("binomial", "poisson", "negative binomial", "Gamma")
    fam = family.family
    if fam in known:
        glm_family = {"negative binomial": "poisson", "Gamma": "Gamma"}.get(fam, fam)
        betas = glm_fit(data.X, data.y, glm_family, data.offset) * np.sqrt(1.346)
    else:
        betas = np.zeros(data.X.shape[1])
    inits = dict(betas=betas, D=np.ones(nRE) if diag_D else np.eye(nRE))
    if data.X_zi is not None:
        y0 = (data.y1() == 0).astype(float)
        inits["gammas"] = glm_fit(data.X_zi, y0, "binomial", data.offset_zi)
        if fam in ("zero-inflated poisson", "zero-inflated negative binomial", "hurdle poisson",
                   "hurdle negative binomial"):
            inits["betas"] = glm_fit(data.X, data.y, "poisson")
    if family.n_phis:
        inits["phis"] = np.zeros(family.n_phis)
    return inits


# ----------------------------------------------------------------------------------------
# numeric derivatives, nearPD
# ----------------------------------------------------------------------------------------
def fd_vec(x, f, eps=DBL_EPS ** 0.25):
    x = np.asarray(x, float)
    n = len(x)
    res = np.zeros((n, n))
    ex = np.maximum(np.abs(x), 1.0)
    f0 = np.asarray(f(x), float)
    for i in range(n):
        x1 = x.copy()
        x1[i] = x[i] + eps * ex[i]
        res[:, i] = (np.asarray(f(x1), float) - f0) / (x1[i] - x[i])
    return 0.5 * (res + res.T)


def nearPD(M, eig_tol=1e-6, conv_tol=1e-7, posd_tol=1e-8, maxits=100):
    M = np.asarray(M, float)
    n = M.shape[1]
    inorm = lambda a: np.max(np.sum(np.abs(a), axis=1))
    U = np.zeros((n, n))
    X = M.copy()
    it, converged = 0, False
    while it < maxits and not converged:
        Y = X
        T = Y - U
        d, Q = np.linalg.eigh(Y)
        d, Q = d[::-1], Q[:, ::-1]                       # eigen(): decreasing
        p = d > eig_tol * d[0]
        QQ = Q[:, p]
        X = QQ @ np.diag(d[p]) @ QQ.T
        U = X - T
        X = (X + X.T) / 2
        conv = inorm(Y - X) / inorm(Y)
        it += 1
        converged = conv <= conv_tol
    X = (X + X.T) / 2
    d, Q = np.linalg.eigh(X)
    d, Q = d[::-1], Q[:, ::-1]
    Eps = posd_tol * abs(d[0])
    if d[n - 1] < Eps:
        d = np.where(d < Eps, Eps, d)
        o_diag = np.diag(X).copy()
        X = Q @ (d[:, None] * Q.T)
        Dd = np.sqrt(np.maximum(Eps, o_diag) / np.diag(X))
        X = Dd[:, None] * X * Dd[None, :]
    return (X + X.T) / 2


# ----------------------------------------------------------------------------------------
# EM pieces (N x K vectorised, as the reference)
# ----------------------------------------------------------------------------------------
def _etas(data, gh, betas, gammas):
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
    return eta_y, eta_zi


def e_step(data, family, gh, betas, D, phis, gammas):
    """R/mixed_fit.R:120-159: p_by, post_b2 (weighted), lgLik and the eta_zi of this iteration."""
    n, nRE = data.n, data.nRE
    eta_y, eta_zi = _etas(data, gh, betas, gammas)
    log_Lik, _ = family.log_dens(data.y, eta_y, phis, eta_zi)
    log_p_yb = agq.rowsum(log_Lik, data)
    K = log_p_yb.shape[1]
    log_p_b = agq.dmvnorm_log(gh.b, np.zeros(nRE), D if nRE > 1 else D.reshape(1, 1)[0]).reshape(n, K)
    log_p_yb_b = log_p_yb + log_p_b
    log_p_y = logsumexp(log_p_yb_b + np.log(gh.wGH)[None, :], axis=1)
    p_by = np.exp(log_p_yb_b - log_p_y[:, None])
    b2 = gh.b2.reshape(n, K, nRE * nRE)
    post_b2 = np.einsum("ik,ikd,k->id", p_by, b2, gh.wGH)
    w = data.weights
    if w is not None:
        post_b2 = w[:, None] * post_b2
    lp = log_p_y + gh.log_dets
    if w is not None:
        lp = w * lp
    lgLik = float(np.sum(lp[np.isfinite(lp)]))
    return p_by, post_b2, lgLik, eta_zi


def _col_scores(Xm, z, p_by, wGH, data):
    sc = np.zeros(Xm.shape[1])
    for l in range(Xm.shape[1]):
        cc = agq.rowsum(Xm[:, [l]] * z, data)
        if data.weights is not None:
            cc = data.weights[:, None] * cc
        sc[l] = np.nansum((cc * p_by) @ wGH)
    return sc


def score_betas(betas, data, family, gh, phis, eta_zi, p_by, penalty=None):
    """R/Fit_Funs.R:295-365 (the penalty term :360-364 added by the wrapper below)."""
    out = _score_betas(betas, data, family, gh, phis, eta_zi, p_by)
    return out if penalty is None else out + penalty.score(betas)


def _score_betas(betas, data, family, gh, phis, eta_zi, p_by):
    canonical, user_defined = agq.family_flags(family)
    eta_y = (data.X @ betas)[:, None] + gh.Ztb
    if data.offset is not None:
        eta_y = eta_y + data.offset[:, None]
    mu_y = family.linkinv(eta_y)
    N = agq.binomial_N(data, family)
    if user_defined:
        z = family.score_eta_fun(data.y, mu_y, phis, eta_zi)
        return -_col_scores(data.X, z, p_by, gh.wGH, data)
    if canonical:
        mu_n = mu_y * N[:, None] if N is not None else mu_y
        sc = _col_scores(data.X, mu_n, p_by, gh.wGH, data)
        y1 = data.y1()
        if data.weights is None:
            Xty = data.X.T @ y1
        else:
            idx = np.repeat(np.arange(data.n), np.diff(data.row_ptr))
            Xty = data.X.T @ (y1 * data.weights[idx])
        return -Xty + sc
    var = family.variance(mu_y)
    deriv = family.mu_eta(eta_y)
    y1 = data.y1()[:, None]
    z = (y1 - (N[:, None] * mu_y if N is not None else mu_y)) * deriv / var
    return -_col_scores(data.X, z, p_by, gh.wGH, data)


def score_phis(phis, data, family, gh, betas, eta_zi, p_by):
    """R/Fit_Funs.R:367-396 (families with a score_phis_fun)."""
    eta_y = (data.X @ betas)[:, None] + gh.Ztb
    if data.offset is not None:
        eta_y = eta_y + data.offset[:, None]
    mu_y = family.linkinv(eta_y)
    z = family.score_phis_fun(data.y, mu_y, phis, eta_zi)
    cc = (agq.rowsum(z, data) * p_by) @ gh.wGH
    # the reference omits the cluster weights in this branch (Fit_Funs.R:392-394); kept as is
    return np.array([-np.nansum(cc)])


def score_gammas(gammas, data, family, gh, betas, phis, p_by):
    """R/Fit_Funs.R:398-427."""
    eta_y, eta_zi = _etas(data, gh, betas, gammas)
    mu_y = family.linkinv(eta_y)
    z = family.score_eta_zi_fun(data.y, mu_y, phis, eta_zi)
    return -_col_scores(data.X_zi, z, p_by, gh.wGH, data)


# ----------------------------------------------------------------------------------------
# mixed_fit
# ----------------------------------------------------------------------------------------
def _pack(betas, D, phis, gammas, diag_D, for_params=False):
    if for_params:       # Params[it, ] (R/mixed_fit.R:118-119): D itself, lower triangle column-major
        q = D.shape[0]
        Dv = np.diag(D) if diag_D else np.array([D[r, c] for c in range(q) for r in range(c, q)])
    else:
        Dv = np.log(np.diag(D)) if diag_D else agq.chol_transf(D)
    parts = [betas, Dv]
    if phis is not None:
        parts.append(phis)
    if gammas is not None:
        parts.append(gammas)
    return np.concatenate([np.atleast_1d(np.asarray(p, float)) for p in parts])


def mixed_fit(data, family, inits=None, control=None, diag_D=False, mode_finder="newton", hessian=True,
              trace=None, penalized=False):
    """R/mixed_fit.R.  mode_finder: "bfgs" is the reference's optim() inside find_modes, "newton"
    the engine's algorithm; both reach the same modes to ~1e-5 / 1e-10."""
    con = dict(CONTROL_DEFAULTS)
    con.update(control or {})
    # R/mixed_model.R:196-208, R/mixed_fit.R:64-69
    if penalized is True:
        penalized = dict(pen_mu=0.0, pen_sigma=1.0, pen_df=3.0)
    pen = agq.Penalty(data.X.shape[1], **penalized) if penalized else None
    nRE = data.nRE
    nAGQ = con["nAGQ"] or (11 if nRE < 3 else 7)
    inits = initial_values(data, family, diag_D) if inits is None else inits
    betas = np.asarray(inits["betas"], float).copy()
    D = np.asarray(inits["D"], float)
    D = np.diag(D) if D.ndim == 1 else D.copy()
    phis = None if inits.get("phis") is None else np.asarray(inits["phis"], float).copy()
    gammas = None if inits.get("gammas") is None else np.asarray(inits["gammas"], float).copy()
    n = data.n
    post_modes = np.zeros((n, nRE))
    numer_deriv_vec = fd_vec if con["numeric_deriv"] == "fd" else agq.cd_vec
    ghfun = lambda pm: agq.GHfun(pm, data, family, betas, np.linalg.inv(D), phis, gammas, nAGQ, nRE,
                                 mode_finder=mode_finder)
    converged = False
    iter_EM = con["iter_EM"]
    if iter_EM > 0:
        update_GH = set(range(0, iter_EM + 1, con["update_GH_every"]))
        Params, lgLik = [], []
        gh = ghfun(post_modes)
        post_modes = gh.post_modes
        for it in range(1, iter_EM + 1):
            if it in update_GH:
                gh = ghfun(post_modes)
                post_modes = gh.post_modes
            Params.append(_pack(betas, D, phis, gammas, diag_D, for_params=True))
            p_by, post_b2, ll, eta_zi = e_step(data, family, gh, betas, D, phis, gammas)
            if pen is not None:
                ll = ll + pen.log_dens(betas)             # R/mixed_fit.R:156-159
            lgLik.append(ll)
            if trace is not None:
                trace.append(("EM", it, ll, betas.copy()))
            if it > 4 and lgLik[-1] > lgLik[-2]:
                th1, th2 = Params[-2], Params[-1]
                check1 = np.max(np.abs(th2 - th1) / (np.abs(th1) + con["tol1"])) < con["tol2"]
                check2 = (lgLik[-1] - lgLik[-2]) < con["tol3"] * (abs(lgLik[-2]) + con["tol3"])
                if check1 or check2:
                    converged = True
                    break
            Dn = np.nanmean(post_b2, axis=0).reshape(nRE, nRE)
            D = 0.5 * (Dn + Dn.T)
            if diag_D:
                D = np.diag(np.diag(D))
            if phis is not None:
                f = lambda ph: score_phis(ph, data, family, gh, betas, eta_zi, p_by)
                H = nearPD(numer_deriv_vec(phis, f))
                phis = phis - np.linalg.solve(H, f(phis))
            f = lambda bt: score_betas(bt, data, family, gh, phis, eta_zi, p_by, pen)
            H = nearPD(numer_deriv_vec(betas, f))
            betas = betas - np.linalg.solve(H, f(betas))
            if gammas is not None:
                f = lambda gm: score_gammas(gm, data, family, gh, betas, phis, p_by)
                H = nearPD(numer_deriv_vec(gammas, f))
                gammas = gammas - np.linalg.solve(H, f(gammas))
            if np.any(np.abs(betas[1:]) > con["max_coef_value"]) or \
                    (gammas is not None and np.any(np.abs(gammas) > con["max_coef_value"])):
                raise FloatingPointError("A large coefficient value has been detected during the optimization.")
    tht = _pack(betas, D, phis, gammas, diag_D)
    p, p_zi = data.X.shape[1], (0 if data.X_zi is None else data.X_zi.shape[1])
    n_phis = family.n_phis
    nD = nRE if diag_D else nRE * (nRE + 1) // 2

    def unpack(t):
        bt, Dp, ph, gm = agq.split_thetas(t, p, nRE, n_phis, p_zi, diag_D)
        Dm = np.diag(np.exp(Dp)) if diag_D else agq.chol_transf(Dp)[0]
        return bt.copy(), Dm, (None if ph is None else ph.copy()), (None if gm is None else gm.copy())

    n_outer = 0
    if not converged and con["iter_qN_outer"] > 0:
        parscale = np.concatenate([np.full(p, con["parscale_betas"]), np.full(nD, con["parscale_D"]),
                                   np.full(n_phis, con["parscale_phis"]), np.full(p_zi, con["parscale_gammas"])])
        iter_qN = con["iter_qN"]
        for it in range(1, con["iter_qN_outer"] + 1):
            n_outer = it
            gh = ghfun(post_modes)
            opt = agq.vmmin(tht, lambda t: agq.logLik_mixed(t, data, family, gh, diag_D, penalty=pen),
                            lambda t: agq.score_mixed(t, data, family, gh, diag_D, penalty=pen),
                            maxit=iter_qN, reltol=con["tol3"], parscale=parscale)
            tht = opt["par"]
            betas, D, phis, gammas = unpack(tht)
            post_modes = gh.post_modes
            if trace is not None:
                trace.append(("qN", it, -opt["value"], betas.copy()))
            if opt["convergence"] == 0:
                converged = True
                break
            iter_qN += con["iter_qN_incr"]
    tht = _pack(betas, D, phis, gammas, diag_D)
    gh = ghfun(post_modes)
    logLik = -agq.logLik_mixed(tht, data, family, gh, diag_D, penalty=pen)
    out = dict(coefficients=betas, D=D, phis=phis, gammas=gammas, post_modes=gh.post_modes,
               post_vars=gh.post_vars, logLik=float(logLik), converged=converged, thetas=tht,
               n_outer=n_outer)
    if hessian:
        out["Hessian"] = agq.cd_vec(tht, lambda t: agq.score_mixed(t, data, family, gh, diag_D, penalty=pen))
    return out
