"""Seeded small data sets, one per accelerated family / link / random-effects shape,
shared by the GPU parity tests and by tests/golden/make_golden.py (which freezes the
oracle's outputs on them as committed fixtures)."""
import numpy as np


def design(rng, n, ni_mean, q_y, zi=False, q_zi=0):
    sizes = 1 + rng.poisson(ni_mean - 1, size=n)
    ids = np.repeat(np.arange(n), sizes)
    N = len(ids)
    x = rng.random(N)
    g = rng.integers(0, 2, n).astype(float)[ids]
    d = dict(id=ids, X=np.column_stack([np.ones(N), x, g]), Z=np.column_stack([np.ones(N), x, x * x - 0.3])[:, :q_y])
    if zi:
        d["X_zi"] = np.column_stack([np.ones(N), g])
        if q_zi:
            d["Z_zi"] = np.ones((N, 1))
    return d, N


def response(fam, rng, N, eta):
    mu = np.exp(np.clip(eta, -3, 3))
    if fam in ("poisson",):
        return rng.poisson(mu).astype(float)
    if fam in ("negative.binomial",):
        return rng.negative_binomial(2.0, 2.0 / (2.0 + mu)).astype(float)
    if fam in ("zi.poisson", "zi.negative.binomial"):
        y = rng.poisson(mu).astype(float)
        y[rng.random(N) < 0.3] = 0.0
        return y
    if fam in ("hurdle.poisson", "hurdle.negative.binomial"):
        y = (1 + rng.poisson(mu)).astype(float)
        y[rng.random(N) < 0.3] = 0.0
        return y
    if fam == "binomial":
        return (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(float)
    if fam == "binomial2":
        trials = rng.integers(1, 7, N)
        s = rng.binomial(trials, 1 / (1 + np.exp(-eta)))
        return np.column_stack([s, trials - s]).astype(float)
    if fam == "zi.binomial":
        trials = rng.integers(1, 6, N)
        s = rng.binomial(trials, 1 / (1 + np.exp(-eta)))
        s[rng.random(N) < 0.25] = 0
        return np.column_stack([s, trials - s]).astype(float)
    if fam == "hurdle.lognormal":
        y = np.exp(eta + 0.5 * rng.standard_normal(N))
        y[rng.random(N) < 0.3] = 0.0
        return y
    if fam == "beta.fam":
        m = 1 / (1 + np.exp(-eta))
        return np.clip(rng.beta(m * 5, (1 - m) * 5), 1e-4, 1 - 1e-4)
    if fam == "hurdle.beta.fam":
        m = 1 / (1 + np.exp(-eta))
        y = np.clip(rng.beta(m * 5, (1 - m) * 5), 1e-4, 1 - 1e-4)
        y[rng.random(N) < 0.3] = 0.0
        return y
    if fam == "Gamma.fam":
        return rng.gamma(3.0, mu / 3.0) + 1e-3
    if fam == "censored.normal":
        v = eta + 0.7 * rng.standard_normal(N)
        hi, lo = np.quantile(v, 0.8), np.quantile(v, 0.1)
        ind = np.where(v > hi, 2.0, np.where(v < lo, 1.0, 0.0))
        return np.column_stack([np.clip(v, lo, hi), ind])
    if fam == "students.t":
        return eta + 0.6 * rng.standard_t(4.0, N)
    if fam in ("beta.binomial", "beta.binomial2"):
        m = 1 / (1 + np.exp(-eta))
        pr = rng.beta(m * 3.0, (1 - m) * 3.0)
        if fam == "beta.binomial":
            return (rng.random(N) < pr).astype(float)
        trials = rng.integers(1, 7, N)
        s = rng.binomial(trials, pr)
        return np.column_stack([s, trials - s]).astype(float)
    raise ValueError(fam)


CASES = [
    # family, link, q_y, q_zi, phis
    ("poisson", None, 1, 0, None),
    ("poisson", None, 3, 0, None),
    ("binomial", "logit", 2, 0, None),
    ("binomial2", "logit", 2, 0, None),
    ("binomial", "probit", 1, 0, None),
    ("binomial", "cloglog", 2, 0, None),
    ("negative.binomial", None, 1, 0, [0.4]),
    ("negative.binomial", None, 3, 0, [0.9]),
    ("zi.poisson", None, 2, 0, None),
    ("zi.poisson", None, 2, 1, None),
    ("zi.negative.binomial", None, 1, 0, [0.5]),
    ("zi.negative.binomial", None, 2, 1, [0.5]),
    ("zi.binomial", None, 1, 1, None),
    ("hurdle.poisson", None, 2, 0, None),
    ("hurdle.poisson", None, 2, 1, None),
    ("hurdle.negative.binomial", None, 2, 1, [0.6]),
    ("hurdle.negative.binomial", None, 1, 0, [0.6]),
    ("hurdle.lognormal", None, 1, 1, [-0.3]),
    ("hurdle.lognormal", None, 2, 0, [-0.3]),
    ("beta.fam", None, 2, 0, [1.5]),
    ("hurdle.beta.fam", None, 1, 1, [1.5]),
    ("Gamma.fam", None, 2, 0, [1.0]),
    ("censored.normal", None, 2, 0, [-0.3]),
    ("students.t", None, 2, 0, [-0.4]),
    ("beta.binomial", None, 1, 0, [1.1]),
    ("beta.binomial2", None, 2, 0, [1.1]),
    ("beta.binomial", "cloglog", 2, 0, [1.1]),
    ("beta.binomial2", "cloglog", 1, 0, [0.6]),
]


def build_case(fam, link, q_y, q_zi, phis, n=36):
    """Returns (fam_name, data dict, theta dict) of one CASES row (seeded)."""
    rng = np.random.default_rng(sum(map(ord, fam)) * 100 + 10 * q_y + q_zi)
    fam_name = {"binomial2": "binomial", "beta.binomial2": "beta.binomial"}.get(fam, fam)
    zi = fam_name.startswith("zi.") or fam_name.startswith("hurdle.")
    d, N = design(rng, n, 6, q_y, zi=zi, q_zi=q_zi)
    q = q_y + q_zi
    A = rng.standard_normal((q, q)) * 0.3
    D = A @ A.T + 0.3 * np.eye(q)
    betas = np.array([0.3, 0.5, -0.4])
    if fam in ("beta.fam", "hurdle.beta.fam", "binomial", "binomial2", "zi.binomial", "beta.binomial", "beta.binomial2"):
        betas = np.array([-0.2, 0.6, -0.3])
    b = rng.standard_normal((n, q)) @ np.linalg.cholesky(D).T
    eta = d["X"] @ betas + np.sum(d["Z"] * b[d["id"], :q_y], axis=1)
    d["y"] = response(fam, rng, N, eta)
    th = dict(betas=betas, D=D, phis=None if phis is None else np.array(phis),
              gammas=np.array([-0.8, 0.4]) if zi else None, nAGQ=7 if q >= 3 else 9,
              df=4.0 if fam == "students.t" else None)
    return fam_name, d, th


def case_id(fam, link, q_y, q_zi, phis):
    return f"{fam}-{link or 'default'}-qy{q_y}-qzi{q_zi}"
