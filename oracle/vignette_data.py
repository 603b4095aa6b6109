"""Regenerates the simulated data sets of the reference's vignettes, whose
rendered outputs (docs/articles/*.html) are the only known-answer values the
reference publishes for the hot path.
TEST INFRASTRUCTURE ONLY.

  binary_data()         vignettes/GLMMadaptive.Rmd:82-107 == vignettes/Methods.Rmd:27-52
  hurdle_lognormal_data()  vignettes/ZeroInflated_and_TwoPart_Models.Rmd:143-178
  hurdle_count_data()   vignettes/ZeroInflated_and_TwoPart_Models.Rmd:313-351
"""
import numpy as np
from scipy import stats
from scipy.special import expit
from .r_rng import RRng


def _design(rng, n, K, t_max):
    # time = c(replicate(n, c(0, sort(runif(K - 1, 0, t_max)))))
    time = np.empty(n * K)
    for i in range(n):
        time[i * K] = 0.0
        time[i * K + 1:(i + 1) * K] = np.sort(rng.runif(K - 1, 0.0, t_max))
    ids = np.repeat(np.arange(n), K)
    sex = np.repeat((np.arange(n) >= n // 2).astype(float), K)   # gl(2, n/2): first half "male"
    X = np.column_stack([np.ones(n * K), sex, time, sex * time])  # ~ sex * time
    Z = np.column_stack([np.ones(n * K), time])                   # ~ time
    return ids, time, sex, X, Z


def binary_data():
    rng = RRng(1234)
    n, K, t_max = 100, 8, 15
    ids, time, sex, X, Z = _design(rng, n, K, t_max)
    betas = np.array([-2.13, -0.25, 0.24, -0.05])
    D11, D22 = 0.48, 0.1
    b = np.column_stack([rng.rnorm(n, sd=np.sqrt(D11)), rng.rnorm(n, sd=np.sqrt(D22))])
    eta_y = X @ betas + np.sum(Z * b[ids], axis=1)
    y = rng.rbinom1(n * K, expit(eta_y))
    return dict(id=ids, time=time, sex=sex, X=X, Z=Z, y=y)


def hurdle_lognormal_data():
    rng = RRng(1234)
    n, K, t_max = 100, 8, 5
    ids, time, sex, X, Z = _design(rng, n, K, t_max)
    X_zi = np.column_stack([np.ones(n * K), sex])
    Z_zi = np.ones((n * K, 1))
    betas = np.array([-2.13, -0.25, 0.24, -0.05])
    sigma = 0.5
    gammas = np.array([-1.5, 0.5])
    D11, D22, D33 = 0.5, 0.1, 0.4
    b = np.column_stack([rng.rnorm(n, sd=np.sqrt(D11)), rng.rnorm(n, sd=np.sqrt(D22)),
                         rng.rnorm(n, sd=np.sqrt(D33))])
    eta_y = X @ betas + np.sum(Z * b[ids, :2], axis=1)
    eta_zi = X_zi @ gammas + Z_zi[:, 0] * b[ids, 2]
    y = np.exp(rng.rnorm(n * K, mean=eta_y, sd=sigma))
    zero = rng.rbinom1(n * K, expit(eta_zi)).astype(bool)
    y[zero] = 0.0
    return dict(id=ids, time=time, sex=sex, X=X, Z=Z, X_zi=X_zi, Z_zi=Z_zi, y=y)


def hurdle_count_data():
    rng = RRng(123)
    n, K, t_max = 100, 8, 5
    ids, time, sex, X, Z = _design(rng, n, K, t_max)
    X_zi = np.column_stack([np.ones(n * K), sex])
    Z_zi = np.ones((n * K, 1))
    betas = np.array([1.5, 0.05, 0.05, -0.03])
    shape = 2.0
    gammas = np.array([-1.5, 0.5])
    D11, D22, D33 = 0.5, 0.1, 0.4
    b = np.column_stack([rng.rnorm(n, sd=np.sqrt(D11)), rng.rnorm(n, sd=np.sqrt(D22)),
                         rng.rnorm(n, sd=np.sqrt(D33))])
    eta_y = X @ betas + np.sum(Z * b[ids, :2], axis=1)
    eta_zi = X_zi @ gammas + Z_zi[:, 0] * b[ids, 2]
    mu = np.exp(eta_y)
    prob = shape / (shape + mu)                     # pnbinom(., mu =, size =)
    lower = stats.nbinom.cdf(0, shape, prob)
    u = rng.runif(n * K, lower, 1.0)
    y = stats.nbinom.ppf(u, shape, prob).astype(float)
    zero = rng.rbinom1(n * K, expit(eta_zi)).astype(bool)
    y[zero] = 0.0
    return dict(id=ids, time=time, sex=sex, X=X, Z=Z, X_zi=X_zi, Z_zi=Z_zi, y=y)


def multcomp_binary_data():
    """vignettes/Multiple_Comparisons.Rmd:30-54: binary outcome at K = 4 categorical time points,
    300 subjects, random intercepts (rnorm and rbinom(., 1, .) only)."""
    rng = RRng(1234)
    n, K = 300, 4
    ids = np.repeat(np.arange(n), K)
    timef = np.tile(np.arange(K), n)                               # gl(K, 1, n * K)
    sex = np.repeat((np.arange(n) >= n // 2).astype(float), K)     # gl(2, n / 2): first half "male"
    T = np.column_stack([(timef == k).astype(float) for k in range(1, K)])
    X_gen = np.column_stack([np.ones(n * K), sex, T, sex[:, None] * T])       # ~ sex * time
    betas = np.array([-2.13, 1.0] + [1.2, -1.2] * (K - 1))
    b = rng.rnorm(n, sd=1.0)
    y = rng.rbinom1(n * K, expit(X_gen @ betas + b[ids]))
    X_main = np.column_stack([np.ones(n * K), sex, T])                        # ~ sex + time
    return dict(id=ids, sex=sex, time=timef, X=X_main, X_interaction=X_gen, Z=np.ones((n * K, 1)), y=y)


# ---- data sets that need R's rpois / rgamma / rnbinom / rbeta (oracle/r_rng.py: RScalarRng) ----
def poisson_data():
    """vignettes/GLMMadaptive.Rmd:201-226 (gm1, gm2): rpois(n * K, exp(eta_y))."""
    from .r_rng import RScalarRng
    rng = RScalarRng(1234)
    n, K, t_max = 100, 8, 15
    ids, time, sex, X, Z = _design(rng, n, K, t_max)
    betas = np.array([2.13, -0.25, 0.24, -0.05])
    b = rng.rnorm(n, sd=np.sqrt(0.48))
    y = rng.rpois(n * K, np.exp(X @ betas + b[ids]))
    return dict(id=ids, time=time, sex=sex, X=X, Z=Z[:, :1], y=y)


def zi_count_data():
    """vignettes/ZeroInflated_and_TwoPart_Models.Rmd:43-79 (fm1, fm2, gm1):
    rnbinom(n * K, size = 2, mu = exp(eta_y)) with extra zeros by rbinom(., 1, plogis(eta_zi))."""
    from .r_rng import RScalarRng
    rng = RScalarRng(1234)
    n, K, t_max = 100, 8, 5
    ids, time, sex, X, Z = _design(rng, n, K, t_max)
    X_zi = np.column_stack([np.ones(n * K), sex])
    Z_zi = np.ones((n * K, 1))
    betas = np.array([1.5, 0.05, 0.05, -0.03])
    gammas = np.array([-1.5, 0.5])
    b = np.column_stack([rng.rnorm(n, sd=np.sqrt(0.5)), rng.rnorm(n, sd=np.sqrt(0.4))])
    eta_y = X @ betas + b[ids, 0]
    eta_zi = X_zi @ gammas + b[ids, 1]
    y = rng.rnbinom_mu(n * K, 2.0, np.exp(eta_y))
    y[rng.rbinom1(n * K, expit(eta_zi)).astype(bool)] = 0.0
    return dict(id=ids, time=time, sex=sex, X=X, Z=Z[:, :1], X_zi=X_zi, Z_zi=Z_zi, y=y)


def custom_negbin_data():
    """vignettes/Custom_Models.Rmd:69-75 (fm1, fm2): expand.grid(f1 = 1:3, f2 = A:B, g = 1:30,
    rep = 1:15) (first factor fastest), rnbinom(mu = mu, size = 0.5); y ~ f1 * f2, ~ 1 | g.
    mixed_model() sorts the rows by the grouping factor (R/mixed_model.R:50: stable order)."""
    from .r_rng import RScalarRng
    rng = RScalarRng(101)
    N = 3 * 2 * 30 * 15
    idx = np.arange(N)
    f1, f2, g = idx % 3 + 1, (idx // 3) % 2 + 1, (idx // 6) % 30
    mu = 5.0 * (-4 + f1 + 4 * f2)
    y = rng.rnbinom_mu(N, 0.5, mu)
    X = np.column_stack([np.ones(N), f1 == 2, f1 == 3, f2 == 2, (f1 == 2) & (f2 == 2),
                         (f1 == 3) & (f2 == 2)]).astype(float)
    o = np.argsort(g, kind="stable")
    return dict(id=g[o], X=X[o], Z=np.ones((N, 1)), y=y[o])


def beta_data():
    """vignettes/Custom_Models.Rmd:262-290 (gm): rbeta(n * K, mu * phi, phi * (1 - mu)), squeezed
    into (0, 1)."""
    from .r_rng import RScalarRng
    rng = RScalarRng(1234)
    n, K, t_max = 100, 8, 15
    ids, time, sex, X, Z = _design(rng, n, K, t_max)
    betas = np.array([-2.2, -0.25, 0.24, -0.05])
    phi = 5.0
    b = rng.rnorm(n, sd=np.sqrt(0.9))
    mu = expit(X @ betas + b[ids])
    y = rng.rbeta(n * K, mu * phi, phi * (1.0 - mu))
    y = (y * (n * K - 1) + 0.5) / (n * K)
    return dict(id=ids, time=time, sex=sex, X=X, Z=Z[:, :1], y=y)
