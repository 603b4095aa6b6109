#!/usr/bin/env python
"""Times one fused evaluation of every accelerated family on the same design (n clusters x 8
observations, q = 2 or 2 + 1, nAGQ = 11 / 7) -- finds families whose point code is out of line.
  python scripts/family_sweep.py [n_clusters]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import glmmadaptive_b200 as gb          # noqa: E402
from family_cases import response       # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
rng = np.random.default_rng(0)
ids = np.repeat(np.arange(n), 8)
N = len(ids)
x = rng.random(N)
g = rng.integers(0, 2, n).astype(float)[ids]
X = np.column_stack([np.ones(N), x, g])
Z = np.column_stack([np.ones(N), x])
X_zi = np.column_stack([np.ones(N), g])
Z_zi = np.ones((N, 1))
D2 = np.array([[0.5, 0.1], [0.1, 0.2]])
D3 = np.array([[0.5, 0.1, -0.1], [0.1, 0.2, 0.05], [-0.1, 0.05, 0.4]])
CASES = [("binomial", "logit", None), ("binomial", "probit", None), ("binomial", "cloglog", None), ("poisson", None, None),
         ("negative.binomial", None, [0.7]), ("Gamma.fam", None, [1.0]), ("beta.fam", None, [1.5]),
         ("censored.normal", None, [-0.3]), ("students.t", None, [-0.4]), ("beta.binomial", None, [1.1]),
         ("zi.poisson", None, None), ("zi.negative.binomial", None, [0.5]), ("zi.binomial", None, None),
         ("hurdle.poisson", None, None), ("hurdle.negative.binomial", None, [0.6]), ("hurdle.lognormal", None, [-0.3]),
         ("hurdle.beta.fam", None, [1.5])]
print(f"{n} clusters x 8 obs; ms per fused evaluation (dominant kernels), point evaluations per second")
for fam, link, phis in CASES:
    zi = fam.startswith("zi.") or fam.startswith("hurdle.")
    betas = np.array([-0.2, 0.6, -0.3]) if ("binomial" in fam or "beta" in fam) else np.array([0.3, 0.5, -0.4])
    b = rng.standard_normal((n, 3)) @ np.linalg.cholesky(D3).T
    eta = X @ betas + np.sum(Z * b[ids, :2], axis=1)
    y = response("zi.binomial" if fam == "zi.binomial" else fam, rng, N, eta)
    kw = dict(X_zi=X_zi, Z_zi=Z_zi) if zi else {}
    q = 3 if zi else 2
    eng = gb.Engine(y, X, Z, ids, fam, link=link, nAGQ=7 if zi else 11, df=4.0 if fam == "students.t" else None, **kw)
    D = D3 if zi else D2
    gam = np.array([-0.8, 0.4]) if zi else None
    ph = None if phis is None else np.array(phis)
    t0 = time.perf_counter()
    st = eng.find_modes(betas, np.linalg.inv(D), ph, gam, warm=False)
    t_modes = time.perf_counter() - t0
    ms_eval, ms_kernel = eng.time_eval(betas, D, ph, gam, True, iters=5)
    r = eng.eval(betas, D, ph, gam)
    print(f"{fam:26s} {link or '':8s} q={q} K={eng.K:4d}  eval {ms_eval:8.3f} ms  kernel {ms_kernel:8.3f} ms  "
          f"{N * eng.K / ms_kernel / 1e6:7.1f} Gpt/s  modes {t_modes * 1e3:7.1f} ms (not conv {st['n_not_converged']})  nonfinite {r['n_nonfinite']}")
    eng.close()
