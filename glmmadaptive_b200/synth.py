"""Synthetic clustered data for the BASELINE.json configurations.
numpy default_rng(seed) (PCG64); x_ij ~ U(0,1); X = [1, x, g] with a cluster-level
g_i ~ Bernoulli(0.5); Z = [1, x] (q_y = 2) or [1]; b_i ~ N(0, D)."""
import numpy as np

CONFIGS = {
    # name: family, q_y, q_zi, nAGQ, betas, D, phis, gammas, default n, default n_i
    "C1": dict(family="binomial", q_y=1, q_zi=0, nAGQ=11, betas=[-0.5, 1.0, -0.5], D=[[1.0]],
               phis=None, gammas=None, n=100, ni=5, seed=1),
    "C2": dict(family="poisson", q_y=2, q_zi=0, nAGQ=11, betas=[0.5, 0.3, -0.2],
               D=[[0.5, 0.1], [0.1, 0.2]], phis=None, gammas=None, n=100_000, ni=10, seed=2),
    "C3": dict(family="negative.binomial", q_y=2, q_zi=0, nAGQ=11, betas=[0.5, 0.3, -0.2],
               D=[[0.5, 0.1], [0.1, 0.2]], phis=[np.log(2.0)], gammas=None, n=1_000_000, ni=8, seed=3),
    "C4": dict(family="zi.poisson", q_y=2, q_zi=1, nAGQ=7, betas=[0.5, 0.3, -0.2],
               D=[[0.5, 0.1, -0.1], [0.1, 0.2, 0.05], [-0.1, 0.05, 0.4]], phis=None,
               gammas=[-1.5, 0.5], n=1_000_000, ni=8, seed=4),
    "C5a": dict(family="censored.normal", q_y=2, q_zi=0, nAGQ=11, betas=[1.0, 0.5, -0.3],
                D=[[0.5, 0.1], [0.1, 0.2]], phis=[np.log(0.7)], gammas=None, n=5_000_000, ni=8, seed=5),
    "C5b": dict(family="beta.fam", q_y=2, q_zi=0, nAGQ=11, betas=[0.5, 0.3, -0.2],
                D=[[0.5, 0.1], [0.1, 0.2]], phis=[np.log(5.0)], gammas=None, n=5_000_000, ni=8, seed=6),
}


def make(name, n=None, ni=None, seed=None, ragged=False, family=None):
    """Returns (data dict for Engine(**data) minus nAGQ, theta dict)."""
    cfg = dict(CONFIGS[name])
    if family is not None:
        cfg["family"] = family
    n = cfg["n"] if n is None else int(n)
    ni = cfg["ni"] if ni is None else int(ni)
    rng = np.random.default_rng(cfg["seed"] if seed is None else seed)
    if ragged:
        sizes = 1 + rng.poisson(max(ni - 1, 0), size=n)
    else:
        sizes = np.full(n, ni, dtype=np.int64)
    N = int(sizes.sum())
    ids = np.repeat(np.arange(n, dtype=np.int64), sizes)
    x = rng.random(N)
    g = rng.integers(0, 2, size=n).astype(float)
    X = np.column_stack([np.ones(N), x, g[ids]])
    q_y, q_zi = cfg["q_y"], cfg["q_zi"]
    Z = np.column_stack([np.ones(N), x])[:, :q_y]
    D = np.array(cfg["D"], float)
    q = q_y + q_zi
    Lc = np.linalg.cholesky(D)
    b = rng.standard_normal((n, q)) @ Lc.T
    betas = np.array(cfg["betas"], float)
    eta = X @ betas + np.sum(Z * b[ids, :q_y], axis=1)
    fam = cfg["family"]
    out = dict(X=X, Z=Z, id=ids, family=fam)
    gammas = None
    if q_zi or cfg["gammas"] is not None:
        gammas = np.array(cfg["gammas"], float)
        X_zi = np.column_stack([np.ones(N), g[ids]])
        out["X_zi"] = X_zi
        eta_zi = X_zi @ gammas
        if q_zi:
            Z_zi = np.ones((N, 1))
            out["Z_zi"] = Z_zi
            eta_zi = eta_zi + b[ids, q_y]
    phis = None if cfg["phis"] is None else np.array(cfg["phis"], float)
    if fam == "binomial":
        y = (rng.random(N) < 1.0 / (1.0 + np.exp(-eta))).astype(float)
    elif fam == "poisson":
        y = rng.poisson(np.exp(eta)).astype(float)
    elif fam == "negative.binomial":
        size = np.exp(phis[0])
        y = rng.negative_binomial(size, size / (size + np.exp(eta))).astype(float)
    elif fam == "zi.poisson":
        y = rng.poisson(np.exp(eta)).astype(float)
        y[rng.random(N) < 1.0 / (1.0 + np.exp(-eta_zi))] = 0.0
    elif fam == "censored.normal":
        sigma = np.exp(phis[0])
        v = eta + sigma * rng.standard_normal(N)
        hi, lo = np.quantile(v, 0.8), np.quantile(v, 0.1)
        ind = np.where(v > hi, 2.0, np.where(v < lo, 1.0, 0.0))
        y = np.column_stack([np.clip(v, lo, hi), ind])
    elif fam == "beta.fam":
        phi = np.exp(phis[0])
        mu = 1.0 / (1.0 + np.exp(-eta))
        v = rng.beta(mu * phi, (1.0 - mu) * phi)
        y = (v * (N - 1) + 0.5) / N
    else:
        raise ValueError(fam)
    out["y"] = y
    theta = dict(betas=betas, D=D, phis=phis, gammas=gammas, nAGQ=cfg["nAGQ"])
    return out, theta
