"""Host-side mirror of the reference's fitting entry points over the CUDA engine:

  mixed_model()   R/mixed_model.R:1-340   (design matrices in, fitted object out; the formula /
                                           model.frame layer of the reference stays with the caller)
  mixed_fit()     R/mixed_fit.R:1-349     (EM phase, quasi-Newton phase, final Hessian)

Everything of size N or n runs on the GPU through libglmma.so (GHfun -> glmma_find_modes, the
E-step -> glmma_em_estep, the M-step scores with p_by held fixed -> glmma_em_score, optim's
fn/gr -> glmma_eval); this module holds the loop logic mixed_fit() keeps in R: the tiny
dense algebra on p-, q- and nparams-sized objects (numeric Hessians of the M-step scores,
nearPD, the Newton updates, optim's BFGS `vmmin`, the convergence rules).  No CPU fallback:
without the CUDA library nothing here can run.

Deviations from the reference, all documented in DESIGN.md:
  * find_modes uses the engine's damped Newton instead of optim(BFGS);
  * the phis M-step uses cluster weights (the reference's score_phis() drops them when the family
    has a score_phis_fun, R/Fit_Funs.R:392-394); identical without weights.
"""
import time

import numpy as np

from ._lib import GlmmaError, ERR_NUMERIC
from .engine import Engine, GHfun, Penalty, chol_transf, logLik_mixed, score_mixed, split_thetas

_EPS = np.finfo(float).eps

# control defaults of mixed_model(), R/mixed_model.R:150-156
CONTROL = dict(iter_EM=30, iter_qN_outer=15, iter_qN=10, iter_qN_incr=10, optimizer="optim",
               optim_method="BFGS", parscale_betas=0.1, parscale_D=0.01, parscale_phis=0.01,
               parscale_gammas=0.01, tol1=1e-4, tol2=1e-5, tol3=1e-8, numeric_deriv="fd", nAGQ=None,
               update_GH_every=10, max_coef_value=12.0, max_phis_value=float(np.exp(10.0)), verbose=False)


# ------------------------------------------------------------------------------------------
# starting values: stats::glm.fit as mixed_model() calls it (R/mixed_model.R:162-194)
# ------------------------------------------------------------------------------------------
class _GlmFamily:
    """initialize / link / variance / deviance pieces of the four stats families used for starts."""

    def __init__(self, name, link=None):
        self.name = name
        # binomial(link = "probit" / "cloglog"): glm.fit is called with the model's own family object
        # (R/mixed_model.R:176), so the starting values follow its link
        self.blink = link if (name == "binomial" and link in ("probit", "cloglog")) else None

    def start(self, y, w):
        if self.name == "binomial":
            return (w * y + 0.5) / (w + 1.0)
        return y + 0.1 if self.name == "poisson" else y

    def link(self, mu):
        if self.blink == "probit":
            from scipy.special import ndtri
            return ndtri(mu)
        if self.blink == "cloglog":
            return np.log(-np.log1p(-mu))
        return {"binomial": lambda m: np.log(m / (1 - m)), "poisson": np.log,
                "Gamma": lambda m: 1 / m, "gaussian": lambda m: m}[self.name](mu)

    def inv(self, eta):
        if self.blink == "probit":
            from scipy.special import ndtr
            th = 8.125890664701906            # -qnorm(.Machine$double.eps)
            return ndtr(np.clip(eta, -th, th))
        if self.blink == "cloglog":
            return np.clip(-np.expm1(-np.exp(eta)), _EPS, 1.0 - _EPS)
        if self.name == "binomial":
            return 1 / (1 + np.exp(-np.clip(eta, -30.0, 30.0)))
        if self.name == "poisson":
            return np.maximum(np.exp(eta), _EPS)
        return 1 / eta if self.name == "Gamma" else eta

    def dmu(self, eta):
        if self.blink == "probit":
            return np.maximum(np.exp(-0.5 * eta * eta) / np.sqrt(2.0 * np.pi), _EPS)
        if self.blink == "cloglog":
            e = np.minimum(eta, 700.0)
            return np.maximum(np.exp(e) * np.exp(-np.exp(e)), _EPS)
        if self.name == "binomial":
            e = np.exp(-np.abs(eta))
            return np.maximum(e / (1 + e) ** 2, _EPS)
        if self.name == "poisson":
            return np.maximum(np.exp(eta), _EPS)
        return -1 / eta ** 2 if self.name == "Gamma" else np.ones_like(eta)

    def var(self, mu):
        return {"binomial": mu * (1 - mu), "poisson": mu, "Gamma": mu * mu, "gaussian": np.ones_like(mu)}[self.name]

    def deviance(self, y, mu, w):
        with np.errstate(divide="ignore", invalid="ignore"):
            if self.name == "binomial":
                r = np.where(y > 0, y * np.log(y / mu), 0.0) + np.where(y < 1, (1 - y) * np.log((1 - y) / (1 - mu)), 0.0)
            elif self.name == "poisson":
                r = np.where(y > 0, y * np.log(y / mu), 0.0) - (y - mu)
            elif self.name == "Gamma":
                r = -(np.log(np.where(y == 0, 1.0, y / mu)) - (y - mu) / mu)
            else:
                return float(np.sum(w * (y - mu) ** 2))
        return float(2 * np.sum(w * r))


def glm_fit(X, y, family, offset=None, epsilon=1e-8, maxit=25, link=None):
    """Iteratively reweighted least squares with glm.fit's starts and stopping rule."""
    X = np.asarray(X, float)
    y = np.asarray(y, float)
    fam = _GlmFamily(family, link)
    w = np.ones(len(X))
    if family == "binomial" and y.ndim == 2:
        w = y.sum(axis=1)
        y = np.divide(y[:, 0], w, out=np.zeros(len(X)), where=w > 0)
    off = 0.0 if offset is None else np.asarray(offset, float)
    eta = fam.link(fam.start(y, w))
    mu = fam.inv(eta)
    dev_old = fam.deviance(y, mu, w)
    beta = np.zeros(X.shape[1])
    for _ in range(maxit):
        g = fam.dmu(eta)
        # weighted least squares through the p x p normal equations (Cholesky): the design has
        # millions of rows and a handful of columns, so a QR of X costs seconds where this costs
        # milliseconds; glm.fit's own QR agrees to the conditioning of X'WX
        ww = w * g * g / fam.var(mu)
        Xw = X * ww[:, None]
        beta = np.linalg.solve(Xw.T @ X, Xw.T @ ((eta - off) + (y - mu) / g))
        eta = X @ beta + off
        mu = fam.inv(eta)
        dev = fam.deviance(y, mu, w)
        if abs(dev - dev_old) / (abs(dev) + 0.1) < epsilon:
            break
        dev_old = dev
    return beta


def starting_values(y, X, X_zi, family, nRE, n_phis, diag_D=False, offset=None, offset_zi=None, link=None):
    """R/mixed_model.R:157-194, 251-263: glm coefficients * sqrt(1.346) for the families the
    reference calls `known`, zeros otherwise; identity D; phis = 0; logistic fit of y == 0 for gammas."""
    fam = family
    starts = {"binomial": "binomial", "poisson": "poisson", "negative binomial": "poisson", "Gamma": "Gamma"}
    if fam in starts:
        betas = glm_fit(X, y, starts[fam], offset, link=link) * np.sqrt(1.346)
    else:
        betas = np.zeros(np.shape(X)[1])
    out = dict(betas=betas, D=np.ones(nRE) if diag_D else np.eye(nRE))
    if X_zi is not None:
        y1 = np.asarray(y, float)
        y1 = y1[:, 0] if y1.ndim == 2 else y1
        out["gammas"] = glm_fit(X_zi, (y1 == 0).astype(float), "binomial", offset_zi)
        if fam in ("zero-inflated poisson", "zero-inflated negative binomial", "hurdle poisson",
                   "hurdle negative binomial"):
            out["betas"] = glm_fit(X, y, "poisson")
    if n_phis:
        out["phis"] = np.zeros(n_phis)
    return out


def glm_fit_device(engine, family, design="X", response="y", use_offset=True, epsilon=1e-8, maxit=25):
    """glm.fit on the data the engine holds: every IRLS iteration is one device pass
    (glmma_glm_pass) returning the deviance and the p x p normal equations; works unchanged on a
    ShardedEngine (the pass results are all-reduced)."""
    p = engine.p if design == "X" else engine.p_zi
    beta, dev_old = None, None
    for it in range(maxit + 1):
        out = engine.glm_pass(family, design, response, use_offset, beta)
        dev = out[0]
        if it > 0 and abs(dev - dev_old) / (abs(dev) + 0.1) < epsilon:
            break
        dev_old = dev
        if it == maxit:
            break
        beta = np.linalg.solve(out[1 + p:].reshape(p, p), out[1:1 + p])
    return beta


def starting_values_device(engine, diag_D=False):
    """starting_values() with the glm.fit calls run on the device-resident data."""
    fam = engine.spec.family
    starts = {"binomial": "binomial", "poisson": "poisson", "negative binomial": "poisson", "Gamma": "Gamma"}
    if fam in starts:
        betas = glm_fit_device(engine, starts[fam]) * np.sqrt(1.346)
    else:
        betas = np.zeros(engine.p)
    out = dict(betas=betas, D=np.ones(engine.q) if diag_D else np.eye(engine.q))
    if engine.p_zi:
        out["gammas"] = glm_fit_device(engine, "binomial", design="X_zi", response="zero")
        if fam in ("zero-inflated poisson", "zero-inflated negative binomial", "hurdle poisson",
                   "hurdle negative binomial"):
            out["betas"] = glm_fit_device(engine, "poisson", use_offset=False)
    if engine.n_phis:
        out["phis"] = np.zeros(engine.n_phis)
    return out


# ------------------------------------------------------------------------------------------
# small dense helpers (R/Functions.R:219-262, 343-382; optim's vmmin)
# ------------------------------------------------------------------------------------------
def fd_vec(x, f, eps=_EPS ** 0.25):
    x = np.asarray(x, float)
    f0 = np.asarray(f(x), float)
    J = np.empty((len(x), len(x)))
    for i in range(len(x)):
        xi = x.copy()
        xi[i] += eps * max(abs(x[i]), 1.0)
        J[:, i] = (np.asarray(f(xi), float) - f0) / (xi[i] - x[i])
    return (J + J.T) / 2


def cd_vec(x, f, eps=1e-3):
    x = np.asarray(x, float)
    J = np.empty((len(x), len(x)))
    for i in range(len(x)):
        hi, lo = x.copy(), x.copy()
        h = eps * max(abs(x[i]), 1.0)
        hi[i] += h
        lo[i] -= h
        J[:, i] = (np.asarray(f(hi), float) - np.asarray(f(lo), float)) / (hi[i] - lo[i])
    return (J + J.T) / 2


def _eig_desc(A):
    w, V = np.linalg.eigh(A)
    return w[::-1], V[:, ::-1]


def nearPD(M, eig_tol=1e-6, conv_tol=1e-7, posd_tol=1e-8, maxits=100):
    """Higham's alternating projections as the reference codes it (R/Functions.R:343-382)."""
    M = np.asarray(M, float)
    k = len(M)
    norm = lambda A: np.abs(A).sum(axis=1).max()
    X, U = M.copy(), np.zeros_like(M)
    for _ in range(maxits):
        Y = X
        T = Y - U
        w, V = _eig_desc(Y)
        keep = w > eig_tol * w[0]
        X = (V[:, keep] * w[keep]) @ V[:, keep].T
        U = X - T
        X = (X + X.T) / 2
        if norm(Y - X) / norm(Y) <= conv_tol:
            break
    X = (X + X.T) / 2
    w, V = _eig_desc(X)
    floor = posd_tol * abs(w[0])
    if w[k - 1] < floor:
        d0 = np.diag(X).copy()
        X = (V * np.where(w < floor, floor, w)) @ V.T
        s = np.sqrt(np.maximum(floor, d0) / np.diag(X))
        X = s[:, None] * X * s[None, :]
    return (X + X.T) / 2


def vmmin(par, fn, gr, maxit, reltol, parscale):
    """optim(method = "BFGS"): R's variable-metric minimiser (src/appl/optim.c `vmmin`, restated from
    its published algorithm) on the parscale-d parameters.
    Returns (par, value, convergence, n_fn, n_gr)."""
    ps = np.asarray(parscale, float)
    F = lambda u: float(fn(u * ps))
    G = lambda u: np.asarray(gr(u * ps), float) * ps
    u = np.asarray(par, float) / ps
    k = len(u)
    if maxit <= 0:
        return u * ps, F(u), 0, 0, 0
    fmin = f = F(u)
    if not np.isfinite(f):
        raise FloatingPointError("initial value in 'vmmin' is not finite")
    g = G(u)
    nf = ng = 1
    it, last_reset = 1, 1
    H = np.eye(k)
    while True:
        if last_reset == ng:
            H = np.eye(k)
        u0, g0 = u.copy(), g.copy()
        step_dir = -H @ g
        slope = float(step_dir @ g)
        if slope < 0.0:
            s, ok = 1.0, False
            while True:
                u = u0 + s * step_dir
                unchanged = int(np.sum((10.0 + u0) == (10.0 + u)))
                if unchanged < k:
                    f = F(u)
                    nf += 1
                    ok = np.isfinite(f) and f <= fmin + slope * s * 1e-4
                    if not ok:
                        s *= 0.2
                if unchanged == k or ok:
                    break
            if not (abs(f - fmin) > reltol * (abs(fmin) + reltol)):
                unchanged = k
                fmin = f
            if unchanged < k:
                fmin = f
                g = G(u)
                ng += 1
                it += 1
                dx = s * step_dir
                dg = g - g0
                curv = float(dx @ dg)
                if curv > 0:
                    Hdg = H @ dg
                    H = H + ((1.0 + float(Hdg @ dg) / curv) * np.outer(dx, dx) - np.outer(Hdg, dx) - np.outer(dx, Hdg)) / curv
                else:
                    last_reset = ng
            elif last_reset < ng:
                unchanged = 0
                last_reset = ng
        else:
            unchanged = 0
            if last_reset == ng:
                unchanged = k
            else:
                last_reset = ng
        if it >= maxit:
            break
        if ng - last_reset > 2 * k:
            last_reset = ng
        if unchanged == k and last_reset == ng:
            break
    return u * ps, fmin, (0 if it < maxit else 1), nf, ng


# ------------------------------------------------------------------------------------------
# mixed_fit
# ------------------------------------------------------------------------------------------
def _params_row(betas, D, phis, gammas, diag_D):
    q = len(D)
    Dv = np.diag(D) if diag_D else np.array([D[r, c] for c in range(q) for r in range(c, q)])
    return np.concatenate([np.atleast_1d(v) for v in (betas, Dv, phis, gammas) if v is not None])


def _thetas(betas, D, phis, gammas, diag_D):
    Dv = np.log(np.diag(D)) if diag_D else chol_transf(D)
    return np.concatenate([np.atleast_1d(v) for v in (betas, Dv, phis, gammas) if v is not None])


def mixed_fit(engine, initial_values, control=None, diag_D=False, hessian=True, trace=None, penalized=False):
    """R/mixed_fit.R:62-349 on a device-resident data set.  Returns the list mixed_fit() returns
    (coefficients, phis, D, gammas, post_modes, post_vars, logLik, Hessian, converged) plus timing
    and evaluation counters."""
    con = dict(CONTROL)
    con.update(control or {})
    eng = engine
    q, n = eng.q, eng.n
    betas = np.array(initial_values["betas"], float)
    D = np.array(initial_values["D"], float)
    D = np.diag(D) if D.ndim == 1 else D
    phis = None if initial_values.get("phis") is None else np.array(initial_values["phis"], float)
    gammas = None if initial_values.get("gammas") is None else np.array(initial_values["gammas"], float)
    nderiv = fd_vec if con["numeric_deriv"] == "fd" else cd_vec
    counts = dict(find_modes=0, estep=0, em_score=0, eval=0)
    pen = Penalty.from_arg(penalized, eng.p)        # R/mixed_fit.R:64-69
    t_start = time.perf_counter()

    def new_grid(cold=False):
        counts["find_modes"] += 1
        return GHfun(eng, 0.0 if cold else None, betas, np.linalg.inv(D), phis, gammas)

    def too_large():
        return bool(np.any(np.abs(betas[1:]) > con["max_coef_value"]) or
                    (gammas is not None and np.any(np.abs(gammas) > con["max_coef_value"])))

    last = {}

    def em_sc(key, bt, ph, gm):
        # the reference evaluates the score at the expansion point twice (inside numer_deriv_vec and
        # again for the Newton step); the second call is answered from the first
        sig = tuple(None if v is None else np.asarray(v, float).tobytes() for v in (bt, ph, gm))
        if sig not in last:
            counts["em_score"] += 1
            last[sig] = eng.em_score(bt, ph, gm)
        return -last[sig][key]

    converged, during_EM = False, False
    gh = None
    if con["iter_EM"] > 0:
        refresh = set(range(0, con["iter_EM"] + 1, con["update_GH_every"]))
        rows, lls = [], []
        gh = new_grid(cold=True)
        for it in range(1, con["iter_EM"] + 1):
            if it in refresh:
                gh = new_grid()
            rows.append(_params_row(betas, D, phis, gammas, diag_D))
            es = eng.em_estep(betas, D, phis, gammas)
            counts["estep"] += 1
            last.clear()
            lls.append(es["loglik"] + (0.0 if pen is None else pen.log_dens(betas)))      # R/mixed_fit.R:156-159
            if trace is not None:
                trace.append(("EM", it, es["loglik"], betas.copy()))
            if it > 4 and lls[-1] > lls[-2]:
                c1 = np.max(np.abs(rows[-1] - rows[-2]) / (np.abs(rows[-2]) + con["tol1"])) < con["tol2"]
                c2 = (lls[-1] - lls[-2]) < con["tol3"] * (abs(lls[-2]) + con["tol3"])
                if c1 or c2:
                    converged = during_EM = True
                    break
            # D <- colMeans(post_b2, na.rm = TRUE), symmetrised (R/mixed_fit.R:186-190)
            D = es["S"] / max(n - es["n_nonfinite"], 1)
            D = (D + D.T) / 2
            if diag_D:
                D = np.diag(np.diag(D))
            # one Newton step per block with numerically differentiated M-step scores, p_by fixed
            if phis is not None:
                f = lambda v: em_sc("g_phi", betas, v, gammas)
                phis = phis - np.linalg.solve(nearPD(nderiv(phis, f)), f(phis))
            f = (lambda v: em_sc("g_beta", v, phis, gammas)) if pen is None else \
                (lambda v: em_sc("g_beta", v, phis, gammas) + pen.score(v))          # R/Fit_Funs.R:360-364
            betas = betas - np.linalg.solve(nearPD(nderiv(betas, f)), f(betas))
            if gammas is not None:
                f = lambda v: em_sc("g_gamma", betas, phis, v)
                gammas = gammas - np.linalg.solve(nearPD(nderiv(gammas, f)), f(gammas))
            if too_large():
                raise FloatingPointError("A large coefficient value has been detected during the optimization "
                                         "(R/mixed_fit.R:70-77); re-scale covariates or set iter_EM = 0")
            if eng.spec.family in ("negative binomial", "zero-inflated negative binomial") and \
                    np.exp(phis[0]) > con["max_phis_value"]:
                raise FloatingPointError("shape/size parameter of the negative binomial above max_phis_value")
    t_em = time.perf_counter() - t_start

    tht = _thetas(betas, D, phis, gammas, diag_D)
    n_outer = 0
    if not converged and con["iter_qN_outer"] > 0:
        nD = q if diag_D else eng.npack
        parscale = np.concatenate([np.full(eng.p, con["parscale_betas"]), np.full(nD, con["parscale_D"]),
                                   np.full(eng.n_phis, con["parscale_phis"]), np.full(eng.p_zi, con["parscale_gammas"])])
        iter_qN = con["iter_qN"]

        def fn(t):
            # a trial step of the line search can leave the region where D = U'U is numerically
            # positive definite (or finite); R's logLik_mixed returns a non-finite value there and
            # vmmin shortens the step (accpoint needs a finite f) -- same here
            counts["eval"] += 1
            try:
                with np.errstate(over="ignore", invalid="ignore"):
                    v = logLik_mixed(t, eng, gh, diag_D, penalty=pen)
            except GlmmaError as e:
                if e.code != ERR_NUMERIC:
                    raise
                return np.inf
            return v if np.isfinite(v) else np.inf

        for it in range(1, con["iter_qN_outer"] + 1):
            n_outer = it
            gh = new_grid(cold=(gh is None))
            tht, val, conv, _, _ = vmmin(tht, fn, lambda t: score_mixed(t, eng, gh, diag_D, penalty=pen), iter_qN,
                                         con["tol3"], parscale)
            betas, D, _, phis, gammas = split_thetas(tht, eng, diag_D)
            betas = betas.copy()
            phis = None if phis is None else phis.copy()
            gammas = None if gammas is None else gammas.copy()
            if trace is not None:
                trace.append(("qN", it, -val, betas.copy()))
            if too_large():
                raise FloatingPointError("A large coefficient value has been detected during the optimization")
            if eng.spec.family in ("negative binomial", "zero-inflated negative binomial") and \
                    np.exp(phis[0]) > con["max_phis_value"]:                          # R/mixed_fit.R:300-303
                raise FloatingPointError("shape/size parameter of the negative binomial above max_phis_value")
            if conv == 0:
                converged = True
                break
            iter_qN += con["iter_qN_incr"]
    t_qn = time.perf_counter() - t_start - t_em

    tht = _thetas(betas, D, phis, gammas, diag_D)
    gh = new_grid(cold=(gh is None))
    out = dict(coefficients=betas, phis=phis, D=D, gammas=gammas, post_modes=gh.post_modes, gh=gh,
               logLik=-logLik_mixed(tht, eng, gh, diag_D, penalty=pen), converged=converged,
               converged_during_EM=during_EM,
               thetas=tht, n_outer=n_outer, counts=counts)
    if hessian:
        # Observed information.  The reference differentiates score_mixed numerically (cd_vec,
        # R/mixed_fit.R:324-342: 2 * nparams evaluations); the engine forms the same matrix in one pass
        # (glmma_observed_information), which control["hessian"] = "cd_vec" turns off.
        H = None
        if con.get("hessian", "analytic") == "analytic" and hasattr(eng, "observed_information"):
            try:
                H = eng.observed_information(betas, D, phis, gammas, diag_D)
                if pen is not None:
                    H[:eng.p, :eng.p] += pen.hessian(betas)
                counts["observed_information"] = 1
            except GlmmaError as e:
                if e.code not in (1,):          # GLMMA_ERR_ARG: more parameters than the one-pass kernel takes
                    raise
                H = None
        if H is None:
            H = cd_vec(tht, lambda t: score_mixed(t, eng, gh, diag_D, penalty=pen))
            counts["eval"] += 2 * len(tht)
        out["Hessian"] = H
    out["seconds"] = dict(EM=t_em, quasi_newton=t_qn, total=time.perf_counter() - t_start)
    return out


def mixed_model(y, X, Z, id, family, link=None, X_zi=None, Z_zi=None, offset=None, offset_zi=None,
                weights=None, initial_values=None, control=None, diag_D=False, device=0, hessian=True,
                engine=None, penalized=False, df=None):
    """mixed_model() from design matrices.  Rows must be grouped by `id` (the reference sorts its
    data frame by the grouping factor, R/mixed_model.R:50).  Returns mixed_fit()'s list.
    ``penalized``: False, True (Student's t penalty on the fixed effects, mean 0, scale 1, 3 df) or a
    dict(pen_mu, pen_sigma, pen_df) (R/mixed_model.R:196-208); ``df``: the degrees of freedom of the
    family object for family = "students.t" (students.t(df), R/Fit_Funs.R:1006)."""
    con = dict(control or {})
    own = engine is None
    if own:
        engine = Engine(y, X, Z, id, family, link=link, nAGQ=con.get("nAGQ"), X_zi=X_zi, Z_zi=Z_zi,
                        offset=offset, offset_zi=offset_zi, weights=weights, device=device, df=df)
    try:
        other_link = engine.spec.family == "binomial" and engine.spec.link != "logit"
        if engine.p <= 8 and engine.p_zi <= 8 and hasattr(engine, "glm_pass") and not other_link:
            inits = starting_values_device(engine, diag_D)
        else:
            # binomial(probit / cloglog): glm.fit with the model's own link (R/mixed_model.R:176), on the host
            inits = starting_values(y, X, X_zi, engine.spec.family, engine.q, engine.n_phis, diag_D, offset, offset_zi,
                                    link=engine.spec.link)
        if initial_values:
            inits.update(initial_values)
        fit = mixed_fit(engine, inits, con, diag_D, hessian, penalized=penalized)
        fit["post_vars"] = fit["gh"].post_vars if engine.n <= 100_000 else None
        fit["nAGQ"] = engine.nAGQ
    finally:
        if own:
            engine.close()
    return fit


# ------------------------------------------------------------------------------------------
# predict(type = "subject_specific") point predictions (R/methods.R:976-1012)
# ------------------------------------------------------------------------------------------
def _linkinv(link, eta):
    """family$linkinv with stats::make.link's clamps."""
    if link == "logit":
        e = np.clip(eta, -30.0, 30.0)       # R: .Call(C_logit_linkinv): exp(eta) clamped to [eps, 1/eps]
        t = np.where(eta < -30, _EPS, np.where(eta > 30, 1.0 / _EPS, np.exp(e)))
        return t / (1.0 + t)
    if link == "log":
        return np.maximum(np.exp(eta), _EPS)
    if link == "probit":
        from math import sqrt
        from scipy.special import ndtr
        thresh = 8.125890664701906        # -qnorm(.Machine$double.eps)
        return ndtr(np.clip(eta, -thresh, thresh))
    if link == "cloglog":
        return np.clip(-np.expm1(-np.exp(eta)), _EPS, 1.0 - _EPS)
    return eta


def predict_subject_specific(fit, y, X, Z, id, family, link=None, X_zi=None, Z_zi=None, offset=None,
                             offset_zi=None, type_pred="response", nAGQ=None, device=0, df=None):
    """predict(object, newdata, type = "subject_specific", se.fit = FALSE): empirical Bayes modes of
    the random effects of the clusters in `newdata` at the fitted parameters -- the device mode search
    of find_modes(), R/methods.R:984-997 -- then eta = X beta + rowSums(Z * b_hat[id, ]) (:998-1003),
    the response-scale prediction and, for families with a zero part, the multiplication by
    1 - plogis(eta_zi) (:1004-1010).  The Monte-Carlo standard errors (:1013-1142) stay with the caller."""
    eng = Engine(y, X, Z, id, family, link=link, nAGQ=nAGQ, X_zi=X_zi, Z_zi=Z_zi, offset=offset,
                 offset_zi=offset_zi, device=device, df=df)
    try:
        betas, D = np.asarray(fit["coefficients"], float), np.asarray(fit["D"], float)
        phis, gammas = fit.get("phis"), fit.get("gammas")
        gh = GHfun(eng, 0.0, betas, np.linalg.inv(D), phis, gammas)
        ids = np.asarray(id)
        idx = np.concatenate(([0], np.cumsum(ids[1:] != ids[:-1])))        # 0-based cluster index per row
        X = np.asarray(X, float)
        Zm = np.asarray(Z, float).reshape(len(X), -1)
        eta = X @ betas + np.sum(Zm * gh.post_modes[idx, :eng.q_y], axis=1)
        if offset is not None:
            eta = eta + np.asarray(offset, float)
        pred = eta if type_pred == "link" else _linkinv(eng.spec.link, eta)
        out = dict(post_modes=gh.post_modes, stats=gh.stats, eta=eta)
        if eng.p_zi:
            eta_zi = np.asarray(X_zi, float).reshape(len(X), -1) @ np.asarray(gammas, float)
            if offset_zi is not None:
                eta_zi = eta_zi + np.asarray(offset_zi, float)
            if eng.q_zi:
                eta_zi = eta_zi + np.sum(np.asarray(Z_zi, float).reshape(len(X), -1) * gh.post_modes[idx, eng.q_y:], axis=1)
            if eng.spec.family == "hurdle log-normal":
                pred = np.exp(pred + 0.5 * np.exp(np.asarray(phis, float)[0]) ** 2)
            zi_probs = 1.0 / (1.0 + np.exp(-eta_zi))
            pred = (1.0 - zi_probs) * pred
            out.update(eta_zi=eta_zi, zi_probs=zi_probs)
        out["pred"] = pred
        if eng.n <= 100_000:
            out["post_vars"] = gh.post_vars
        return out
    finally:
        eng.close()
