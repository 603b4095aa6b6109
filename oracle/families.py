"""Family objects of the reference, restated in numpy (TEST INFRASTRUCTURE ONLY).

Every family mirrors the list of closures the reference builds in
R/Fit_Funs.R: ``log_dens(y, eta, mu_fun, phis, eta_zi)`` (returning the
N x K matrix of log densities and, as the R attribute ``"mu_y"``, the mean),
``score_eta_fun``, ``score_phis_fun`` and ``score_eta_zi_fun`` (all taking
``(y, mu, phis, eta_zi)``).  ``eta``/``mu`` are N x K matrices (column k = AGQ
node k), ``y`` is a length-N vector or an N x 2 matrix, ``eta_zi`` is a length-N
vector or an N x K matrix -- exactly the shapes R recycles over.

R's ``stats`` arithmetic (dbinom, plogis, pnorm, dgamma, make.link) is not under
/root/reference; it is substituted here by scipy.special as SURVEY.md section
8(c) lists, with R's link clamps restated.
"""
import numpy as np
from scipy import special as sp

DBL_EPS = np.finfo(float).eps
SQRT_EPS = np.sqrt(DBL_EPS)


# ----------------------------------------------------------------------------
# link functions (R stats::make.link / binomial(); clamps per SURVEY appendix C)
# ----------------------------------------------------------------------------
def logit_linkinv(eta):
    """binomial()$linkinv == C_logit_linkinv: exp(eta) clamped at +-30."""
    eta = np.asarray(eta, float)
    with np.errstate(over="ignore"):
        t = np.where(eta < -30.0, DBL_EPS, np.where(eta > 30.0, 1.0 / DBL_EPS, np.exp(eta)))
    return t / (1.0 + t)


def logit_mu_eta(eta):
    """binomial()$mu.eta == C_logit_mu_eta."""
    eta = np.asarray(eta, float)
    with np.errstate(over="ignore"):
        opexp = 1.0 + np.exp(eta)
        out = np.exp(eta) / (opexp * opexp)
    return np.where(np.abs(eta) > 30.0, DBL_EPS, out)


def log_linkinv(eta):
    """make.link("log")$linkinv = pmax(exp(eta), .Machine$double.eps)."""
    with np.errstate(over="ignore"):
        return np.maximum(np.exp(np.asarray(eta, float)), DBL_EPS)


def identity_linkinv(eta):
    return np.asarray(eta, float)


def probit_linkinv(eta):
    thresh = -sp.ndtri(DBL_EPS)
    return sp.ndtr(np.clip(np.asarray(eta, float), -thresh, thresh))


def probit_mu_eta(eta):
    return np.maximum(np.exp(-0.5 * np.asarray(eta, float) ** 2) / np.sqrt(2 * np.pi), DBL_EPS)


def cloglog_linkinv(eta):
    eta = np.asarray(eta, float)
    with np.errstate(over="ignore"):
        return np.maximum(np.minimum(-np.expm1(-np.exp(eta)), 1 - DBL_EPS), DBL_EPS)


def cloglog_mu_eta(eta):
    eta = np.minimum(np.asarray(eta, float), 700.0)
    return np.maximum(np.exp(eta) * np.exp(-np.exp(eta)), DBL_EPS)


def _col(v):
    """R recycles a length-N vector down the rows of an N x K matrix."""
    v = np.asarray(v, float)
    return v[:, None] if v.ndim == 1 else v


def plogis(x):
    return sp.expit(x)


def plogis_log(x):          # plogis(x, log.p = TRUE)
    return sp.log_expit(x)


def plogis_upper_log(x):    # plogis(x, lower.tail = FALSE, log.p = TRUE)
    return sp.log_expit(-np.asarray(x, float))


def dbinom_log(x, n, p):
    """R's dbinom(x, n, p, log = TRUE) (nmath/dbinom.c, dbinom_raw) restated:
    x == 0 -> n*log(q) with the p < 0.1 branch evaluated as n*log1p(-p);
    x == n -> n*log(p) with the q < 0.1 branch evaluated as n*log1p(-q);
    otherwise the saddle-point form, which equals
    lchoose(n, x) + x log p + (n - x) log q to rounding."""
    x, n, p = np.broadcast_arrays(np.asarray(x, float), np.asarray(n, float), np.asarray(p, float))
    q = 1.0 - p
    with np.errstate(divide="ignore", invalid="ignore"):
        lq = np.where(p < 0.1, np.log1p(-p), np.log(q))
        lp = np.where(q < 0.1, np.log1p(-q), np.log(p))
        gen = (sp.gammaln(n + 1.0) - sp.gammaln(x + 1.0) - sp.gammaln(n - x + 1.0)
               + sp.xlogy(x, p) + sp.xlogy(n - x, q))
    out = np.where(x == 0, n * lq, np.where(x == n, n * lp, gen))
    return np.where(n == 0, 0.0, out)


def dnorm_log(x, mean, sd):
    z = (x - mean) / sd
    return -0.5 * z * z - np.log(sd) - 0.5 * np.log(2 * np.pi)


class Family:
    family = None
    link = None
    has_phis = False
    n_phis = 0
    zi = False          # family has an extra zero part (needs eta_zi)
    y_cols = 1
    score_phis_fun = None
    score_eta_zi_fun = None
    score_eta_fun = None
    variance = None
    mu_eta = None

    def linkinv(self, eta):
        raise NotImplementedError

    # mu_fun in the reference is family$linkinv (R/mixed_model.R:211-215)
    def mu_fun(self, eta):
        return self.linkinv(eta)


# ----------------------------------------------------------------------------
# stats::binomial() / stats::poisson() with the package's own log densities
# ----------------------------------------------------------------------------
class binomial(Family):
    """R/Fit_Funs.R:429-438 (binomial_log_dens); scores by the canonical /
    non-canonical branches of score_mixed (R/Fit_Funs.R:127-167)."""
    family = "binomial"

    def __init__(self, link="logit"):
        self.link = link
        self._inv, self._mu_eta = {
            "logit": (logit_linkinv, logit_mu_eta),
            "probit": (probit_linkinv, probit_mu_eta),
            "cloglog": (cloglog_linkinv, cloglog_mu_eta),
        }[link]

    def linkinv(self, eta):
        return self._inv(eta)

    def mu_eta(self, eta):
        return self._mu_eta(eta)

    def variance(self, mu):
        return mu * (1.0 - mu)

    def log_dens(self, y, eta, phis=None, eta_zi=None):
        mu_y = self.mu_fun(eta)
        y = np.asarray(y, float)
        if y.ndim == 2:
            out = dbinom_log(_colv(y[:, 0], mu_y), _colv(y[:, 0] + y[:, 1], mu_y), mu_y)
        else:
            out = dbinom_log(_colv(y, mu_y), 1.0, mu_y)
        return out, mu_y


def _colv(v, like):
    v = np.asarray(v, float)
    return v[:, None] if np.ndim(like) == 2 else v


class poisson(Family):
    """R/Fit_Funs.R:440-445 (poisson_log_dens)."""
    family = "poisson"
    link = "log"

    def linkinv(self, eta):
        return log_linkinv(eta)

    def mu_eta(self, eta):
        return log_linkinv(eta)

    def variance(self, mu):
        return mu

    def log_dens(self, y, eta, phis=None, eta_zi=None):
        mu_y = self.mu_fun(eta)
        yc = _colv(y, mu_y)
        out = yc * np.log(mu_y) - mu_y - sp.gammaln(yc + 1.0)
        return out, mu_y


# ----------------------------------------------------------------------------
# package families
# ----------------------------------------------------------------------------
class negative_binomial(Family):
    """R/Fit_Funs.R:467-513."""
    family = "negative binomial"
    link = "log"
    has_phis = True
    n_phis = 1

    def linkinv(self, eta):
        return log_linkinv(eta)

    def log_dens(self, y, eta, phis, eta_zi=None):
        phis = np.exp(phis[0])
        mu = self.mu_fun(eta)
        yc = _colv(y, mu)
        log_mu_phis = np.log(mu + phis)
        comp1 = sp.gammaln(yc + phis) - sp.gammaln(phis) - sp.gammaln(yc + 1.0)
        comp2 = phis * np.log(phis) - phis * log_mu_phis
        comp3 = yc * (np.log(mu) - log_mu_phis)
        return comp1 + comp2 + comp3, mu

    def score_eta_fun(self, y, mu, phis, eta_zi=None):
        phis = np.exp(phis[0])
        yc = _colv(y, mu)
        mu_mu_phis = mu / (mu + phis)
        return -phis * mu_mu_phis + yc * (1.0 - mu_mu_phis)

    def score_phis_fun(self, y, mu, phis, eta_zi=None):
        phis = np.exp(phis[0])
        yc = _colv(y, mu)
        mu_phis = mu + phis
        y_phis = yc + phis
        comp1 = np.log(phis) + 1.0 - sp.digamma(phis)
        comp2 = sp.digamma(y_phis)
        comp3 = -np.log(mu_phis) - y_phis / mu_phis
        return (comp1 + comp2 + comp3) * phis


class zi_poisson(Family):
    """R/Fit_Funs.R:515-563."""
    family = "zero-inflated poisson"
    link = "log"
    zi = True

    def linkinv(self, eta):
        return log_linkinv(eta)

    def log_dens(self, y, eta, phis, eta_zi):
        y = np.asarray(y, float)
        ind_y0 = y == 0
        ind_y1 = y > 0
        mu = _col(self.mu_fun(eta))
        with np.errstate(over="ignore"):
            lam = _col(np.exp(eta_zi))
        lam = np.broadcast_to(lam, mu.shape)
        out = np.array(_col(eta), float, copy=True)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            out[ind_y0, :] = np.log(lam[ind_y0, :] + np.exp(-mu[ind_y0, :]))
            y1 = y[ind_y1][:, None]
            out[ind_y1, :] = y1 * np.log(mu[ind_y1, :]) - mu[ind_y1, :] - sp.gammaln(y1 + 1.0)
            out = out - np.log(1.0 + lam)
        return out, mu

    def score_eta_fun(self, y, mu, phis, eta_zi):
        y = np.asarray(y, float)
        ind_y0 = y == 0
        ind_y1 = y > 0
        mu = _col(mu)
        with np.errstate(over="ignore"):
            lam = np.broadcast_to(_col(np.exp(eta_zi)), mu.shape)
        out = np.array(mu, float, copy=True)
        with np.errstate(over="ignore", invalid="ignore"):
            out[ind_y0, :] = -mu[ind_y0, :] / (lam[ind_y0, :] * np.exp(mu[ind_y0, :]) + 1.0)
        out[ind_y1, :] = y[ind_y1][:, None] - mu[ind_y1, :]
        return out

    def score_eta_zi_fun(self, y, mu, phis, eta_zi):
        y = np.asarray(y, float)
        ind_y0 = y == 0
        mu = _col(mu)
        eta_zi = np.broadcast_to(_col(eta_zi), mu.shape)
        with np.errstate(over="ignore", invalid="ignore"):
            lam = np.exp(eta_zi)
            l0 = np.exp(eta_zi[ind_y0, :] + mu[ind_y0, :])
            l0 = np.where(l0 == np.inf, 1e200, l0)
            out = np.array(-lam / (1.0 + lam), float, copy=True)
            out[ind_y0, :] = out[ind_y0, :] + l0 / (l0 + 1.0)
        return out


class zi_negative_binomial(Family):
    """R/Fit_Funs.R:565-640."""
    family = "zero-inflated negative binomial"
    link = "log"
    zi = True
    has_phis = True
    n_phis = 1

    def linkinv(self, eta):
        return log_linkinv(eta)

    def log_dens(self, y, eta, phis, eta_zi):
        phis = np.exp(phis[0])
        y = np.asarray(y, float)
        mu = _col(self.mu_fun(eta))
        yc = y[:, None]
        log_mu_phis = np.log(mu + phis)
        comp1 = sp.gammaln(yc + phis) - sp.gammaln(phis) - sp.gammaln(yc + 1.0)
        comp2 = phis * np.log(phis) - phis * log_mu_phis
        comp3 = yc * (np.log(mu) - log_mu_phis)
        out = np.array(comp1 + comp2 + comp3, float, copy=True)
        ind_y0 = y == 0
        ind_y1 = y > 0
        pis = np.broadcast_to(_col(plogis(eta_zi)), mu.shape)
        with np.errstate(divide="ignore"):
            out[ind_y0, :] = np.log(pis[ind_y0, :] + (1.0 - pis[ind_y0, :]) * np.exp(out[ind_y0, :]))
            out[ind_y1, :] = np.log(1.0 - pis[ind_y1, :]) + out[ind_y1, :]
        return out, mu

    def score_eta_fun(self, y, mu, phis, eta_zi):
        phis = np.exp(phis[0])
        y = np.asarray(y, float)
        mu = _col(mu)
        mu_mu_phis = mu / (mu + phis)
        out = np.array(-phis * mu_mu_phis + y[:, None] * (1.0 - mu_mu_phis), float, copy=True)
        ind_y0 = y == 0
        with np.errstate(over="ignore"):
            lam = np.exp(np.broadcast_to(_col(eta_zi), mu.shape)[ind_y0, :])
        mu0 = mu[ind_y0, :]
        t = phis / (phis + mu0)
        den = (lam + t ** phis) * (mu0 + phis) ** 2
        out[ind_y0, :] = -phis ** 2 * t ** (phis - 1.0) * mu0 / den
        return out

    def score_eta_zi_fun(self, y, mu, phis, eta_zi):
        phis = np.exp(phis[0])
        y = np.asarray(y, float)
        ind_y0 = y == 0
        ind_y1 = y > 0
        mu = _col(mu)
        with np.errstate(over="ignore"):
            lam = np.broadcast_to(_col(np.exp(eta_zi)), mu.shape)
        out = np.array(mu, float, copy=True)
        out[ind_y1, :] = -lam[ind_y1, :] / (1.0 + lam[ind_y1, :])
        t = phis / (phis + mu[ind_y0, :])
        lam0 = lam[ind_y0, :]
        out[ind_y0, :] = lam0 / (lam0 + t ** phis) - lam0 / (1.0 + lam0)
        return out

    def score_phis_fun(self, y, mu, phis, eta_zi):
        phis = np.exp(phis[0])
        y = np.asarray(y, float)
        mu = _col(mu)
        yc = y[:, None]
        mu_phis = mu + phis
        comp1 = sp.digamma(yc + phis) - sp.digamma(phis)
        comp2 = np.log(phis) + 1.0 - np.log(mu_phis) - phis / mu_phis
        comp3 = -yc / mu_phis
        out = np.array((comp1 + comp2 + comp3) * phis, float, copy=True)
        ind_y0 = y == 0
        with np.errstate(over="ignore"):
            lam = np.broadcast_to(_col(np.exp(eta_zi)), mu.shape)
        t = phis / (phis + mu[ind_y0, :])
        t_phis = t ** phis
        out[ind_y0, :] = t_phis * (np.log(t) + 1.0 - t) * phis / (lam[ind_y0, :] + t_phis)
        return out


class _hurdle_zero_part:
    """score_eta_zi_fun shared by all hurdle families, e.g. R/Fit_Funs.R:670-676."""

    def score_eta_zi_fun(self, y, mu, phis, eta_zi):
        y = np.asarray(y, float)
        ind = y > 0
        mu = _col(mu)
        probs = np.broadcast_to(plogis(_col(eta_zi)), mu.shape)
        out = np.array(1.0 - probs, float, copy=True)
        out[ind, :] = -probs[ind, :]
        return out


class hurdle_poisson(_hurdle_zero_part, Family):
    """R/Fit_Funs.R:642-688."""
    family = "hurdle poisson"
    link = "log"
    zi = True

    def linkinv(self, eta):
        return log_linkinv(eta)

    def log_dens(self, y, eta, phis, eta_zi):
        y = np.asarray(y, float)
        ind = y > 0
        eta = _col(eta)
        mu = self.mu_fun(eta)
        eta_zi = np.broadcast_to(_col(eta_zi), eta.shape)
        out = np.array(eta, float, copy=True)
        yi = y[ind][:, None]
        with np.errstate(divide="ignore"):
            out[ind, :] = (plogis_upper_log(eta_zi[ind, :]) - mu[ind, :] + yi * eta[ind, :]
                           - np.log(-np.expm1(-mu[ind, :])) - sp.gammaln(yi + 1.0))
        out[~ind, :] = plogis_log(eta_zi[~ind, :])
        return out, mu

    def score_eta_fun(self, y, mu, phis, eta_zi):
        y = np.asarray(y, float)
        ind = y > 0
        mu = _col(mu)
        mu_ind = mu[ind, :]
        out = np.array(mu, float, copy=True)
        out[~ind, :] = 0.0
        out[ind, :] = -mu_ind + y[ind][:, None] + (np.exp(-mu_ind) * mu_ind) / np.expm1(-mu_ind)
        return out


class hurdle_negative_binomial(_hurdle_zero_part, Family):
    """R/Fit_Funs.R:690-764."""
    family = "hurdle negative binomial"
    link = "log"
    zi = True
    has_phis = True
    n_phis = 1

    def linkinv(self, eta):
        return log_linkinv(eta)

    def log_dens(self, y, eta, phis, eta_zi):
        phis = np.exp(phis[0])
        y = np.asarray(y, float)
        ind = y > 0
        eta = _col(eta)
        mu = self.mu_fun(eta)
        yc = y[:, None]
        log_mu_phis = np.log(mu + phis)
        eta_zi = np.broadcast_to(_col(eta_zi), eta.shape)
        out = np.array(eta, float, copy=True)
        comp1 = sp.gammaln(yc + phis) - sp.gammaln(phis) - sp.gammaln(yc + 1.0)
        comp2 = phis * np.log(phis) - phis * log_mu_phis
        comp3 = yc * np.log(mu) - yc * log_mu_phis
        log_g = comp1 + comp2 + comp3
        with np.errstate(divide="ignore"):
            comp4 = np.log(1.0 - (1.0 + mu / phis) ** (-phis))
        out[ind, :] = plogis_upper_log(eta_zi[ind, :]) + log_g[ind, :] - comp4[ind, :]
        out[~ind, :] = plogis_log(eta_zi[~ind, :])
        return out, mu

    def score_eta_fun(self, y, mu, phis, eta_zi):
        phis = np.exp(phis[0])
        y = np.asarray(y, float)
        ind = y > 0
        mu = _col(mu)
        yc = y[:, None]
        mu_phis = mu + phis
        comp2 = -phis / mu_phis
        comp3 = yc / mu - yc / mu_phis
        k = 1.0 + mu / phis
        with np.errstate(divide="ignore", invalid="ignore"):
            comp4 = k ** (-phis - 1.0) / (1.0 - k ** (-phis))
        out = np.array((comp2 + comp3 - comp4) * mu, float, copy=True)
        out[~ind, :] = 0.0
        return out

    def score_phis_fun(self, y, mu, phis, eta_zi):
        y = np.asarray(y, float)
        ind_y0 = y == 0
        phis = np.exp(phis[0])
        mu = _col(mu)
        yc = y[:, None]
        mu_phis = mu + phis
        comp1 = sp.digamma(yc + phis) - sp.digamma(phis)
        comp2 = np.log(phis) + 1.0 - np.log(mu_phis) - phis / mu_phis
        comp3 = -yc / mu_phis
        k = mu / phis
        k1 = 1.0 + k
        with np.errstate(divide="ignore", invalid="ignore"):
            comp4 = k1 ** (-phis) * (k / k1 - np.log(k1)) / (1.0 - k1 ** (-phis))
        out = np.array((comp1 + comp2 + comp3 + comp4) * phis, float, copy=True)
        out[ind_y0, :] = 0.0
        return out


class zi_binomial(Family):
    """R/Fit_Funs.R:766-827 (score_phis_fun = NULL, no phis)."""
    family = "zero-inflated binomial"
    link = "logit"
    zi = True

    def linkinv(self, eta):
        return logit_linkinv(eta)

    @staticmethod
    def _yN(y):
        y = np.asarray(y, float)
        if y.ndim == 2:
            return y[:, 0], y[:, 0] + y[:, 1]
        return y, np.ones_like(y)

    def log_dens(self, y, eta, phis, eta_zi):
        mu = _col(self.mu_fun(eta))
        y1, N = self._yN(y)
        out = np.array(dbinom_log(y1[:, None], N[:, None], mu), float, copy=True)
        ind_y0 = y1 == 0
        ind_y1 = y1 > 0
        pis = np.broadcast_to(_col(plogis(eta_zi)), mu.shape)
        with np.errstate(divide="ignore"):
            out[ind_y0, :] = np.log(pis[ind_y0, :] + (1.0 - pis[ind_y0, :]) * np.exp(out[ind_y0, :]))
            out[ind_y1, :] = np.log(1.0 - pis[ind_y1, :]) + out[ind_y1, :]
        return out, mu

    def score_eta_fun(self, y, mu, phis, eta_zi):
        mu = _col(mu)
        y1, N = self._yN(y)
        out = np.array(y1[:, None] * (1.0 - mu) - (N - y1)[:, None] * mu, float, copy=True)
        ind_y0 = y1 == 0
        eta_zi = np.broadcast_to(_col(eta_zi), mu.shape)
        pis = plogis(eta_zi[ind_y0, :])
        mu0 = mu[ind_y0, :]
        N0 = N[ind_y0][:, None]
        pis1 = 1.0 - pis
        den = pis + pis1 * (1.0 - mu0) ** N0
        out[ind_y0, :] = -(N0 * mu0 * pis1 * (1.0 - mu0) ** N0) / den
        return out

    def score_eta_zi_fun(self, y, mu, phis, eta_zi):
        y1, N = self._yN(y)
        ind_y0 = y1 == 0
        ind_y1 = y1 > 0
        mu = _col(mu)
        pis = np.broadcast_to(_col(plogis(eta_zi)), mu.shape)
        out = np.array(mu, float, copy=True)
        out[ind_y1, :] = -pis[ind_y1, :]
        mu0 = mu[ind_y0, :]
        N0 = N[ind_y0][:, None]
        pis1 = 1.0 - pis[ind_y0, :]
        FF = (1.0 - mu0) ** N0
        den = pis[ind_y0, :] + pis1 * FF
        out[ind_y0, :] = pis[ind_y0, :] * pis1 * (1.0 - FF) / den
        return out


class hurdle_lognormal(_hurdle_zero_part, Family):
    """R/Fit_Funs.R:829-885 (attr mu_y = eta; identity link)."""
    family = "hurdle log-normal"
    link = "identity"
    zi = True
    has_phis = True
    n_phis = 1

    def linkinv(self, eta):
        return identity_linkinv(eta)

    def log_dens(self, y, eta, phis, eta_zi):
        sigma = np.exp(phis[0])
        y = np.asarray(y, float)
        ind = y > 0
        eta = _col(eta)
        eta_zi = np.broadcast_to(_col(eta_zi), eta.shape)
        out = np.array(eta, float, copy=True)
        out[ind, :] = plogis_upper_log(eta_zi[ind, :]) + dnorm_log(np.log(y[ind])[:, None], eta[ind, :], sigma)
        out[~ind, :] = plogis_log(eta_zi[~ind, :])
        return out, eta

    def score_eta_fun(self, y, mu, phis, eta_zi):
        sigma = np.exp(phis[0])
        y = np.asarray(y, float)
        ind = y > 0
        eta = _col(mu)
        out = np.array(eta, float, copy=True)
        out[~ind, :] = 0.0
        out[ind, :] = (np.log(y[ind])[:, None] - eta[ind, :]) / sigma ** 2
        return out

    def score_phis_fun(self, y, mu, phis, eta_zi):
        sigma = np.exp(phis[0])
        y = np.asarray(y, float)
        ind = y > 0
        eta = _col(mu)
        out = np.array(eta, float, copy=True)
        out[~ind, :] = 0.0
        out[ind, :] = -1.0 + (np.log(y[ind])[:, None] - eta[ind, :]) ** 2 / sigma ** 2
        return out


class beta_fam(Family):
    """R/Fit_Funs.R:887-930."""
    family = "beta"
    link = "logit"
    has_phis = True
    n_phis = 1

    def linkinv(self, eta):
        return logit_linkinv(eta)

    def log_dens(self, y, eta, phis, eta_zi=None):
        phi = np.exp(phis[0])
        mu = self.mu_fun(eta)
        yc = _colv(y, mu)
        mu_phi = mu * phi
        comp1 = sp.gammaln(phi) - sp.gammaln(mu_phi)
        comp2 = (mu_phi - 1.0) * np.log(yc) - sp.gammaln(phi - mu_phi)
        comp3 = (phi - mu_phi - 1.0) * np.log(1.0 - yc)
        return comp1 + comp2 + comp3, mu

    def score_eta_fun(self, y, mu, phis, eta_zi=None):
        phi = np.exp(phis[0])
        yc = _colv(y, mu)
        mu_phi = mu * phi
        comp1 = -sp.digamma(mu_phi) * phi
        comp2 = phi * (np.log(yc) + sp.digamma(phi - mu_phi))
        comp3 = -phi * np.log(1.0 - yc)
        mu_eta = mu - mu * mu
        return (comp1 + comp2 + comp3) * mu_eta

    def score_phis_fun(self, y, mu, phis, eta_zi=None):
        phi = np.exp(phis[0])
        yc = _colv(y, mu)
        mu_phi = mu * phi
        mu1 = 1.0 - mu
        comp1 = sp.digamma(phi) - sp.digamma(mu_phi) * mu
        comp2 = mu * np.log(yc) - sp.digamma(phi - mu_phi) * mu1
        comp3 = np.log(1.0 - yc) * mu1
        return (comp1 + comp2 + comp3) * phi


class hurdle_beta_fam(_hurdle_zero_part, Family):
    """R/Fit_Funs.R:932-1004.

    Quirk: the reference's log_dens sets attr "mu_y" to
    *eta* (R/Fit_Funs.R:951), so score_mixed hands eta to score_eta_fun /
    score_phis_fun where they expect mu (find_modes passes the true mu,
    R/Functions.R:34-37).  ``quirk_mu_is_eta`` selects which of the two this
    object reports as "mu_y": False (default) = the mathematically correct mu,
    True = the reference's literal behaviour inside score_mixed."""
    family = "hurdle beta"
    link = "logit"
    zi = True
    has_phis = True
    n_phis = 1

    def __init__(self, quirk_mu_is_eta=False):
        self.quirk_mu_is_eta = quirk_mu_is_eta

    def linkinv(self, eta):
        return logit_linkinv(eta)

    def log_dens(self, y, eta, phis, eta_zi):
        phi = np.exp(phis[0])
        y = np.asarray(y, float)
        ind = y > 0
        eta = _col(eta)
        eta_zi = np.broadcast_to(_col(eta_zi), eta.shape)
        out = np.array(eta, float, copy=True)
        mu = self.mu_fun(eta)
        mu_phi = mu * phi
        yc = y[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            comp1 = sp.gammaln(phi) - sp.gammaln(mu_phi)
            comp2 = (mu_phi - 1.0) * np.log(yc) - sp.gammaln(phi - mu_phi)
            comp3 = (phi - mu_phi - 1.0) * np.log(1.0 - yc)
        out[ind, :] = plogis_upper_log(eta_zi[ind, :]) + comp1[ind, :] + comp2[ind, :] + comp3[ind, :]
        out[~ind, :] = plogis_log(eta_zi[~ind, :])
        return out, (eta if self.quirk_mu_is_eta else mu)

    def score_eta_fun(self, y, mu, phis, eta_zi):
        phi = np.exp(phis[0])
        y = np.asarray(y, float)
        ind = y > 0
        mu = _col(mu)
        mu_phi = mu * phi
        yc = y[:, None]
        out = np.array(mu, float, copy=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            comp1 = -sp.digamma(mu_phi) * phi
            comp2 = phi * (np.log(yc) + sp.digamma(phi - mu_phi))
            comp3 = -phi * np.log(1.0 - yc)
        mu_eta = mu - mu * mu
        out[~ind, :] = 0.0
        out[ind, :] = (comp1[ind, :] + comp2[ind, :] + comp3[ind, :]) * mu_eta[ind, :]
        return out

    def score_phis_fun(self, y, mu, phis, eta_zi):
        phi = np.exp(phis[0])
        y = np.asarray(y, float)
        ind = y > 0
        mu = _col(mu)
        mu_phi = mu * phi
        mu1 = 1.0 - mu
        yc = y[:, None]
        out = np.array(mu, float, copy=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            comp1 = sp.digamma(phi) - sp.digamma(mu_phi) * mu
            comp2 = mu * np.log(yc) - sp.digamma(phi - mu_phi) * mu1
            comp3 = np.log(1.0 - yc) * mu1
        out[ind, :] = (comp1[ind, :] + comp2[ind, :] + comp3[ind, :]) * phi
        out[~ind, :] = 0.0
        return out


class Gamma_fam(Family):
    """R/Fit_Funs.R:1323-1358; dgamma(y, shape = phi, scale = mu/phi, log = TRUE)."""
    family = "Gamma"
    link = "log"
    has_phis = True
    n_phis = 1

    def linkinv(self, eta):
        return log_linkinv(eta)

    def log_dens(self, y, eta, phis, eta_zi=None):
        phi = np.exp(phis[0])
        eta = _col(eta)
        mu_y = self.mu_fun(eta)
        yc = np.asarray(y, float)[:, None]
        scale = mu_y / phi
        out = (phi - 1.0) * np.log(yc) - yc / scale - sp.gammaln(phi) - phi * np.log(scale)
        return out, mu_y

    def score_eta_fun(self, y, mu, phis, eta_zi=None):
        phi = np.exp(phis[0])
        mu = _col(mu)
        yc = np.asarray(y, float)[:, None]
        comp = phi / mu
        return comp * (yc / mu - 1.0) * mu

    def score_phis_fun(self, y, mu, phis, eta_zi=None):
        phi = np.exp(phis[0])
        mu = _col(mu)
        yc = np.asarray(y, float)[:, None]
        comp1 = np.log(yc) - np.log(mu) - yc / mu
        comp2 = np.log(phi) + 1.0 - sp.digamma(phi)
        return (comp1 + comp2) * phi


class censored_normal(Family):
    """R/Fit_Funs.R:1360-1435; y = [value, indicator in {0 none, 1 left, 2 right}]."""
    family = "censored normal"
    link = "identity"
    has_phis = True
    n_phis = 1
    y_cols = 2

    def linkinv(self, eta):
        return identity_linkinv(eta)

    def log_dens(self, y, eta, phis, eta_zi=None):
        sigma = np.exp(phis[0])
        y = np.asarray(y, float)
        ind0, ind1, ind2 = y[:, 1] == 0, y[:, 1] == 1, y[:, 1] == 2
        eta = _col(eta)
        out = np.array(eta, float, copy=True)
        out[ind0, :] = dnorm_log(y[ind0, 0][:, None], eta[ind0, :], sigma)
        out[ind1, :] = sp.log_ndtr((y[ind1, 0][:, None] - eta[ind1, :]) / sigma)
        out[ind2, :] = sp.log_ndtr(-(y[ind2, 0][:, None] - eta[ind2, :]) / sigma)
        return out, eta

    def score_eta_fun(self, y, mu, phis, eta_zi=None):
        sigma = np.exp(phis[0])
        y = np.asarray(y, float)
        ind0, ind1, ind2 = y[:, 1] == 0, y[:, 1] == 1, y[:, 1] == 2
        eta = _col(mu)
        out = np.array(eta, float, copy=True)
        if ind0.any():
            out[ind0, :] = (y[ind0, 0][:, None] - eta[ind0, :]) / sigma ** 2
        if ind1.any():
            tt = (y[ind1, 0][:, None] - eta[ind1, :]) / sigma
            A = sp.ndtr(tt)
            with np.errstate(divide="ignore", invalid="ignore"):
                out[ind1, :] = -np.exp(-0.5 * tt ** 2) / (np.sqrt(2 * np.pi) * sigma * A)
        if ind2.any():
            tt = (y[ind2, 0][:, None] - eta[ind2, :]) / sigma
            P = sp.ndtr(tt)
            A = eta[ind2, :] * P - sigma * np.exp(-0.5 * tt ** 2) / np.sqrt(2 * np.pi)
            B = np.maximum(sp.ndtr(-tt), SQRT_EPS)
            out[ind2, :] = (-A / B + eta[ind2, :] * (1.0 - B) / B) / sigma ** 2
        return out

    def score_phis_fun(self, y, mu, phis, eta_zi=None):
        sigma = np.exp(phis[0])
        y = np.asarray(y, float)
        ind0, ind1, ind2 = y[:, 1] == 0, y[:, 1] == 1, y[:, 1] == 2
        eta = _col(mu)
        out = np.array(eta, float, copy=True)
        if ind0.any():
            out[ind0, :] = -1.0 + (y[ind0, 0][:, None] - eta[ind0, :]) ** 2 / sigma ** 2
        if ind1.any():
            tt = (y[ind1, 0][:, None] - eta[ind1, :]) / sigma
            A = (-tt * np.exp(-0.5 * tt ** 2)) / np.sqrt(2 * np.pi)
            B = np.maximum(sp.ndtr(tt), SQRT_EPS)
            out[ind1, :] = A / B
        if ind2.any():
            tt = (y[ind2, 0][:, None] - eta[ind2, :]) / sigma
            P = sp.ndtr(tt)
            A = sigma ** 2 * P + sigma ** 2 * (-tt * np.exp(-0.5 * tt ** 2)) / np.sqrt(2 * np.pi)
            B = np.maximum(sp.ndtr(-tt), SQRT_EPS)
            out[ind2, :] = (1.0 - B) / B - A / (B * sigma ** 2)
        return out


ALL_FAMILIES = {
    "binomial": binomial,
    "poisson": poisson,
    "negative.binomial": negative_binomial,
    "zi.poisson": zi_poisson,
    "zi.negative.binomial": zi_negative_binomial,
    "zi.binomial": zi_binomial,
    "hurdle.poisson": hurdle_poisson,
    "hurdle.negative.binomial": hurdle_negative_binomial,
    "hurdle.lognormal": hurdle_lognormal,
    "beta.fam": beta_fam,
    "hurdle.beta.fam": hurdle_beta_fam,
    "Gamma.fam": Gamma_fam,
    "censored.normal": censored_normal,
}


class students_t(Family):
    """R/Fit_Funs.R:1006-1041; identity link, df fixed by the family object."""
    family = "Student's-t"
    link = "identity"
    has_phis = True
    n_phis = 1

    def __init__(self, df):
        self.df = float(df)

    def linkinv(self, eta):
        return identity_linkinv(eta)

    def variance(self, mu):
        return np.ones_like(mu)

    def log_dens(self, y, eta, phis, eta_zi=None):
        sigma = np.exp(phis[0])
        yc = _colv(y, eta)
        df = self.df
        t = (yc - eta) / sigma
        out = sp.gammaln((df + 1) / 2) - sp.gammaln(df / 2) - 0.5 * np.log(df * np.pi) \
            - (df + 1) / 2 * np.log1p(t * t / df) - np.log(sigma)        # dt(t, df, log = TRUE) - log(sigma)
        return out, eta

    def score_eta_fun(self, y, mu, phis, eta_zi=None):
        sigma2 = np.exp(phis[0]) ** 2
        y_mu = _colv(y, mu) - mu
        return (y_mu * (self.df + 1) / (self.df * sigma2)) / (1 + y_mu ** 2 / (self.df * sigma2))

    def score_phis_fun(self, y, mu, phis, eta_zi=None):
        sigma = np.exp(phis[0])
        y_mu2_df = (_colv(y, mu) - mu) ** 2 / self.df
        return (self.df + 1) * y_mu2_df * sigma ** -2 / (1 + y_mu2_df / sigma ** 2) - 1


class beta_binomial(Family):
    """R/Fit_Funs.R:1245-1321; link "logit" (default) or "cloglog" (the two the reference's score_eta_fun
    switches on, R/Fit_Funs.R:1288-1290)."""
    family = "beta binomial"
    link = "logit"
    has_phis = True
    n_phis = 1

    def __init__(self, link="logit"):
        if link not in ("logit", "cloglog"):
            raise ValueError("beta.binomial: link must be logit or cloglog")
        self.link = link

    def linkinv(self, eta):
        return logit_linkinv(eta) if self.link == "logit" else cloglog_linkinv(eta)

    def _size_y(self, y, like):
        y = np.asarray(y, float)
        if y.ndim == 2:
            return _colv(y[:, 0] + y[:, 1], like), _colv(y[:, 0], like)
        return _colv(np.ones(len(y)), like), _colv(y, like)

    def log_dens(self, y, eta, phis, eta_zi=None):
        phi = np.exp(phis[0])
        mu_y = self.mu_fun(eta)
        size, x = self._size_y(y, mu_y)
        A, B = phi * mu_y, phi * (1 - mu_y)
        lbeta = lambda a, b: sp.gammaln(a) + sp.gammaln(b) - sp.gammaln(a + b)
        lchoose = sp.gammaln(size + 1) - sp.gammaln(x + 1) - sp.gammaln(size - x + 1)
        return lchoose + lbeta(x + A, size - x + B) - lbeta(A, B), mu_y

    def score_eta_fun(self, y, mu, phis, eta_zi=None):
        phi = np.exp(phis[0])
        size, x = self._size_y(y, mu)
        phi_mu, phi_1mu = phi * mu, phi * (1 - mu)
        comp1 = (sp.digamma(x + phi_mu) - sp.digamma(size - x + phi_1mu)) * phi
        comp2 = (sp.digamma(phi_mu) - sp.digamma(phi_1mu)) * phi
        # mu.eta <- switch(.link, "logit" = mu - mu * mu, "cloglog" = - (1 - mu) * log(1 - mu))
        mu_eta = (mu - mu * mu) if self.link == "logit" else -(1 - mu) * np.log(1 - mu)
        return (comp1 - comp2) * mu_eta

    def score_phis_fun(self, y, mu, phis, eta_zi=None):
        phi = np.exp(phis[0])
        size, x = self._size_y(y, mu)
        mu1 = 1 - mu
        phi_mu, phi_1mu = phi * mu, phi * mu1
        comp1 = sp.digamma(x + phi_mu) * mu + sp.digamma(size - x + phi_1mu) * mu1 - sp.digamma(size + phi)
        comp2 = sp.digamma(phi_mu) * mu + sp.digamma(phi_1mu) * mu1 - sp.digamma(phi)
        return (comp1 - comp2) * phi
