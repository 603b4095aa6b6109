"""Pair tile of the register variant (q = 1) against the one-cluster-per-warp path on the same data."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import glmmadaptive_b200 as gb

fam = sys.argv[1] if len(sys.argv) > 1 else "negative.binomial"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
rng = np.random.default_rng(5)
sizes = rng.integers(1, 9, n)
ids = np.repeat(np.arange(n), sizes)
N = len(ids)
x = rng.random(N)
g = rng.integers(0, 2, n).astype(float)[ids]
X = np.column_stack([np.ones(N), x, g])
Z = np.ones((N, 1))
betas = np.array([0.3, 0.5, -0.4])
D = np.array([[0.6]])
b = rng.normal(0, np.sqrt(0.6), n)
eta = X @ betas + b[ids]
phis = None
if fam == "negative.binomial":
    y = rng.negative_binomial(2.0, 2.0 / (2.0 + np.exp(eta))).astype(float); phis = np.array([np.log(2.0)])
elif fam == "poisson":
    y = rng.poisson(np.exp(eta)).astype(float)
else:
    y = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(float)
w = rng.uniform(0.5, 1.5, n)
res = {}
for tag, env in (("pair", None), ("single", "1")):
    if env: os.environ["GLMMA_NO_PAIR"] = env
    else: os.environ.pop("GLMMA_NO_PAIR", None)
    eng = gb.Engine(y, X, Z, ids, fam, nAGQ=11, weights=w)
    eng.find_modes(betas, np.linalg.inv(D), phis, None, warm=False)
    r = eng.eval_raw(betas, D, phis, None)
    t0 = time.perf_counter()
    for _ in range(20): r = eng.eval_raw(betas, D, phis, None)
    dt = (time.perf_counter() - t0) / 20
    c = eng.eval_contrib(betas, D, phis, None)
    es = eng.em_estep(betas, D, phis, None); ms = eng.em_score(betas + 0.02, phis, None)
    res[tag] = (r, c, es, ms, dt)
    eng.close()
r1, c1, e1, m1, t1 = res["pair"]; r2, c2, e2, m2, t2 = res["single"]
print(fam, "max diff / largest component %.2e" % (np.max(np.abs(r1 - r2)) / np.max(np.abs(r2))), "per component", np.max(np.abs(r1 - r2) / np.maximum(1, np.abs(r2))))
print("contrib ll_i %.2e zbar %.2e S_i %.2e" % (np.max(np.abs(c1["ll_i"] - c2["ll_i"])), np.max(np.abs(c1["zbar"] - c2["zbar"])), np.max(np.abs(c1["S_i"] - c2["S_i"]))))
print("estep ll %.2e  mstep g_beta %.2e" % (abs(e1["loglik"] - e2["loglik"]) / abs(e2["loglik"]), np.max(np.abs(m1["g_beta"] - m2["g_beta"]))))
print("ms per eval: pair %.3f single %.3f" % (t1 * 1e3, t2 * 1e3))
