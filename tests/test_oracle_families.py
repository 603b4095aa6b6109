"""CPU: every family of the oracle (oracle/families.py: log_dens, score_eta_fun, score_phis_fun,
score_eta_zi_fun) against mpmath at 50 digits.

The mpmath side is written from the TEXTBOOK definition of each distribution (not from the R code the
oracle restates): a log density as a function of (eta, phis, eta_zi), and the three scores as mpmath
numerical derivatives of that log density.  So a wrong sign, a dropped term or a lost factor in any of
the reference's hand-derived score formulas (R/Fit_Funs.R:429-1435) -- or in the
oracle's restatement of it -- shows up here.  Tolerance: relative 1e-12 on log densities, 1e-10
(relative to the larger of |score| and the magnitude of the cancelling terms) on scores; the
reference's sqrt(eps) floors and link clamps are kept inactive by the chosen arguments (they are
exercised by tests/test_gpu_edges.py against the oracle).
"""
import mpmath as mp
import numpy as np
import pytest

from oracle import families as F

mp.mp.dps = 50
RTOL_LD = 1e-12
RTOL_SC = 1e-10


def _plogis(x):
    return 1 / (1 + mp.exp(-x))


def _lchoose(n, k):
    return mp.loggamma(n + 1) - mp.loggamma(k + 1) - mp.loggamma(n - k + 1)


def _lbeta(a, b):
    return mp.loggamma(a) + mp.loggamma(b) - mp.loggamma(a + b)


def _lnorm(x, m, s):
    return -mp.log(s) - mp.log(2 * mp.pi) / 2 - ((x - m) / s) ** 2 / 2


def _Phi(x):
    return mp.erfc(-x / mp.sqrt(2)) / 2


def _pois(y, mu):
    return y * mp.log(mu) - mu - mp.loggamma(y + 1)


def _nb(y, mu, phi):
    return (mp.loggamma(y + phi) - mp.loggamma(phi) - mp.loggamma(y + 1) + phi * mp.log(phi / (mu + phi))
            + y * mp.log(mu / (mu + phi)))


def _beta(y, mu, phi):
    return -_lbeta(mu * phi, (1 - mu) * phi) + (mu * phi - 1) * mp.log(y) + ((1 - mu) * phi - 1) * mp.log(1 - y)


# textbook log densities: f(y, eta, phi_unconstrained, eta_zi); y is a scalar or a tuple
def ld_binomial(link):
    inv = {"logit": _plogis, "probit": _Phi, "cloglog": lambda e: 1 - mp.exp(-mp.exp(e))}[link]

    def f(y, eta, ph, ez):
        s, fl = y
        p = inv(eta)
        return _lchoose(s + fl, s) + s * mp.log(p) + fl * mp.log(1 - p)
    return f


def ld_poisson(y, eta, ph, ez):
    return _pois(y, mp.exp(eta))


def ld_negbin(y, eta, ph, ez):
    return _nb(y, mp.exp(eta), mp.exp(ph))


def ld_zi_poisson(y, eta, ph, ez):
    pi, mu = _plogis(ez), mp.exp(eta)
    return mp.log(pi + (1 - pi) * mp.exp(-mu)) if y == 0 else mp.log(1 - pi) + _pois(y, mu)


def ld_zi_negbin(y, eta, ph, ez):
    pi, mu, phi = _plogis(ez), mp.exp(eta), mp.exp(ph)
    return mp.log(pi + (1 - pi) * mp.exp(_nb(0, mu, phi))) if y == 0 else mp.log(1 - pi) + _nb(y, mu, phi)


def ld_zi_binomial(y, eta, ph, ez):
    s, fl = y
    pi, p = _plogis(ez), _plogis(eta)
    b = _lchoose(s + fl, s) + s * mp.log(p) + fl * mp.log(1 - p)
    return mp.log(pi + (1 - pi) * mp.exp(b)) if s == 0 else mp.log(1 - pi) + b


def ld_hurdle_poisson(y, eta, ph, ez):
    pi, mu = _plogis(ez), mp.exp(eta)
    return mp.log(pi) if y == 0 else mp.log(1 - pi) + _pois(y, mu) - mp.log(1 - mp.exp(-mu))


def ld_hurdle_negbin(y, eta, ph, ez):
    pi, mu, phi = _plogis(ez), mp.exp(eta), mp.exp(ph)
    return mp.log(pi) if y == 0 else mp.log(1 - pi) + _nb(y, mu, phi) - mp.log(1 - mp.exp(_nb(0, mu, phi)))


def ld_hurdle_lognormal(y, eta, ph, ez):
    pi = _plogis(ez)
    return mp.log(pi) if y == 0 else mp.log(1 - pi) + _lnorm(mp.log(y), eta, mp.exp(ph))


def ld_beta(y, eta, ph, ez):
    return _beta(y, _plogis(eta), mp.exp(ph))


def ld_hurdle_beta(y, eta, ph, ez):
    pi = _plogis(ez)
    return mp.log(pi) if y == 0 else mp.log(1 - pi) + _beta(y, _plogis(eta), mp.exp(ph))


def ld_gamma(y, eta, ph, ez):
    shape, mu = mp.exp(ph), mp.exp(eta)
    scale = mu / shape
    return -mp.loggamma(shape) - shape * mp.log(scale) + (shape - 1) * mp.log(y) - y / scale


def ld_censored_normal(y, eta, ph, ez):
    v, ind = y
    s = mp.exp(ph)
    if ind == 0:
        return _lnorm(v, eta, s)
    t = (v - eta) / s
    return mp.log(_Phi(t)) if ind == 1 else mp.log(_Phi(-t))


def ld_students_t(df):
    def f(y, eta, ph, ez):
        s = mp.exp(ph)
        t = (y - eta) / s
        return (mp.loggamma((df + 1) / 2) - mp.loggamma(df / 2) - mp.log(df * mp.pi) / 2
                - (df + 1) / 2 * mp.log(1 + t * t / df) - mp.log(s))
    return f


def ld_beta_binomial(y, eta, ph, ez):
    s, fl = y
    phi, mu = mp.exp(ph), _plogis(eta)
    a, b = phi * mu, phi * (1 - mu)
    return _lchoose(s + fl, s) + _lbeta(s + a, fl + b) - _lbeta(a, b)


def ld_beta_binomial_cloglog(y, eta, ph, ez):
    s, fl = y
    phi, mu = mp.exp(ph), 1 - mp.exp(-mp.exp(eta))
    a, b = phi * mu, phi * (1 - mu)
    return _lchoose(s + fl, s) + _lbeta(s + a, fl + b) - _lbeta(a, b)


TWO_COL = [(0.0, 5.0), (2.0, 3.0), (5.0, 0.0), (1.0, 0.0), (0.0, 1.0), (7.0, 13.0)]
COUNTS = [0.0, 1.0, 2.0, 7.0, 30.0, 250.0]
UNIT = [0.003, 0.2, 0.5, 0.77, 0.995]
CENS = [(0.3, 0.0), (-1.2, 0.0), (0.5, 1.0), (-0.7, 1.0), (1.4, 2.0), (-0.2, 2.0), (2.5, 2.0)]

#        name, oracle family, mpmath log density, y values, eta values, phis values, eta_zi values
CASES = [
    ("binomial-logit", lambda: F.binomial("logit"), ld_binomial("logit"), TWO_COL, [-4.0, -0.3, 0.0, 1.7, 6.0], [None], [None]),
    ("binomial-probit", lambda: F.binomial("probit"), ld_binomial("probit"), TWO_COL, [-2.5, -0.3, 0.0, 1.1, 3.0], [None], [None]),
    ("binomial-cloglog", lambda: F.binomial("cloglog"), ld_binomial("cloglog"), TWO_COL, [-4.0, -0.3, 0.0, 0.9, 2.0], [None], [None]),
    ("poisson", F.poisson, ld_poisson, COUNTS, [-3.0, -0.2, 0.4, 2.5, 5.5], [None], [None]),
    ("negative.binomial", F.negative_binomial, ld_negbin, COUNTS, [-3.0, -0.2, 0.4, 2.5, 5.5], [-1.2, 0.0, 0.7, 3.0], [None]),
    ("zi.poisson", F.zi_poisson, ld_zi_poisson, COUNTS, [-3.0, -0.2, 0.4, 2.5], [None], [-3.0, -0.4, 0.0, 2.0]),
    ("zi.negative.binomial", F.zi_negative_binomial, ld_zi_negbin, COUNTS, [-3.0, -0.2, 0.4, 2.5], [-1.2, 0.7, 2.0], [-3.0, -0.4, 2.0]),
    ("zi.binomial", F.zi_binomial, ld_zi_binomial, TWO_COL, [-4.0, -0.3, 0.0, 1.7], [None], [-3.0, -0.4, 0.0, 2.0]),
    ("hurdle.poisson", F.hurdle_poisson, ld_hurdle_poisson, COUNTS, [-3.0, -0.2, 0.4, 2.5], [None], [-3.0, -0.4, 0.0, 2.0]),
    ("hurdle.negative.binomial", F.hurdle_negative_binomial, ld_hurdle_negbin, COUNTS, [-3.0, -0.2, 0.4, 2.5], [-1.2, 0.7, 2.0], [-3.0, -0.4, 2.0]),
    ("hurdle.lognormal", F.hurdle_lognormal, ld_hurdle_lognormal, [0.0, 0.02, 0.9, 4.0, 300.0], [-3.0, -0.2, 0.4, 2.5], [-1.0, 0.0, 0.8], [-3.0, -0.4, 2.0]),
    ("beta.fam", F.beta_fam, ld_beta, UNIT, [-3.0, -0.4, 0.0, 1.3, 3.0], [-0.5, 0.9, 1.7, 3.5], [None]),
    ("hurdle.beta.fam", lambda: F.hurdle_beta_fam(quirk_mu_is_eta=False), ld_hurdle_beta, [0.0] + UNIT, [-3.0, -0.4, 1.3], [-0.5, 1.7, 3.5], [-3.0, -0.4, 2.0]),
    ("Gamma.fam", F.Gamma_fam, ld_gamma, [0.01, 0.5, 1.0, 7.5, 120.0], [-2.0, -0.2, 0.4, 2.5], [-1.0, 0.0, 1.5, 3.0], [None]),
    ("censored.normal", F.censored_normal, ld_censored_normal, CENS, [-1.5, -0.1, 0.6, 2.0], [-0.7, 0.0, 0.5], [None]),
    ("students.t", lambda: F.students_t(4.0), ld_students_t(4), [-3.0, 0.1, 2.2, 15.0], [-1.5, 0.0, 0.6, 2.0], [-0.7, 0.0, 0.5], [None]),
    ("beta.binomial", F.beta_binomial, ld_beta_binomial, TWO_COL, [-3.0, -0.3, 0.0, 1.7], [-1.0, 0.0, 1.5, 3.0], [None]),
    ("beta.binomial-cloglog", lambda: F.beta_binomial("cloglog"), ld_beta_binomial_cloglog, TWO_COL, [-2.0, -0.3, 0.0, 0.9], [-1.0, 0.0, 1.5, 3.0], [None]),
]


def _y_array(y):
    return np.array([y], float) if not isinstance(y, tuple) else np.array([list(y)], float)


def _y_mp(y):
    return tuple(mp.mpf(v) for v in y) if isinstance(y, tuple) else mp.mpf(y)


def _binomial_like_scores(fam, name, y, eta, mu):
    """R/Fit_Funs.R:127-167: binomial / poisson have no score_eta_fun; score_mixed uses y - N mu
    (canonical) or (y - N mu) mu.eta / var (other links)."""
    if name == "poisson":
        return y[0] - mu
    s, N = y[0, 0], y[0, 0] + y[0, 1]
    if fam.link == "logit":
        return s - N * mu
    return (s - N * mu) * fam.mu_eta(eta) / fam.variance(mu)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_family_against_mpmath(case):
    name, make, ld, ys, etas, phs, ezs = case
    fam = make()
    n_checked = 0
    for y in ys:
        ya, ym = _y_array(y), _y_mp(y)
        if name in ("binomial-logit", "binomial-probit", "binomial-cloglog") and ya[0].sum() == 1:
            ya_call = ya[:, 0]                      # the 0/1 vector form of a binary outcome
        elif name in ("zi.binomial", "beta.binomial", "beta.binomial-cloglog") and ya[0].sum() == 1:
            ya_call = ya[:, 0]
        else:
            ya_call = ya
        for eta in etas:
            for ph in phs:
                for ez in ezs:
                    phis = None if ph is None else np.array([ph])
                    eta_a = np.array([[eta]])
                    ez_a = None if ez is None else np.array([[ez]])
                    out, mu = fam.log_dens(ya_call, eta_a, phis, ez_a)
                    e_m, p_m, z_m = mp.mpf(eta), (None if ph is None else mp.mpf(ph)), (None if ez is None else mp.mpf(ez))
                    ref = ld(ym, e_m, p_m, z_m)
                    got = float(np.asarray(out).ravel()[0])
                    assert abs(got - float(ref)) <= RTOL_LD * max(1.0, abs(float(ref))), (name, y, eta, ph, ez, got, float(ref))
                    if name == "censored.normal" and y[1] > 0 and abs((y[0] - eta) / np.exp(ph)) > 4.5:
                        continue        # the reference's sqrt(eps) floor on the tail probability is active
                    # d/d eta
                    ref_e = mp.diff(lambda e: ld(ym, e, p_m, z_m), e_m)
                    if fam.score_eta_fun is not None and name not in ("poisson",) and not name.startswith("binomial"):
                        got_e = float(np.asarray(fam.score_eta_fun(ya_call, mu, phis, ez_a)).ravel()[0])
                    else:
                        got_e = float(np.ravel(_binomial_like_scores(fam, name, ya if ya.ndim == 2 and ya.shape[1] == 2 else ya, eta_a, mu))[0])
                    scale = max(1.0, abs(float(ref_e)), abs(float(ref)))
                    assert abs(got_e - float(ref_e)) <= RTOL_SC * scale, (name, "eta", y, eta, ph, ez, got_e, float(ref_e))
                    if ph is not None:
                        ref_p = mp.diff(lambda p: ld(ym, e_m, p, z_m), p_m)
                        got_p = float(np.asarray(fam.score_phis_fun(ya_call, mu, phis, ez_a)).ravel()[0])
                        assert abs(got_p - float(ref_p)) <= RTOL_SC * max(scale, abs(float(ref_p))), (name, "phis", y, eta, ph, ez, got_p, float(ref_p))
                    if ez is not None:
                        ref_z = mp.diff(lambda z: ld(ym, e_m, p_m, z), z_m)
                        got_z = float(np.asarray(fam.score_eta_zi_fun(ya_call, mu, phis, ez_a)).ravel()[0])
                        assert abs(got_z - float(ref_z)) <= RTOL_SC * max(1.0, abs(float(ref_z))), (name, "eta_zi", y, eta, ph, ez, got_z, float(ref_z))
                    n_checked += 1
    assert n_checked >= 20
