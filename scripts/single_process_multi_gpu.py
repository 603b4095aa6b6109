#!/usr/bin/env python
"""One process driving several GPUs through glmma_create_multi (what an R session would do):
BASELINE config 3 (1e6 clusters), evaluations per second through glmma_eval (host theta in, host
result out) and a whole mixed_model() fit, for 1, 2, 4, 8 devices of the box.
  python scripts/single_process_multi_gpu.py [n_clusters]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import glmmadaptive_b200 as gb          # noqa: E402
from glmmadaptive_b200 import synth     # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d, th = synth.make("C3", n=n, seed=None)
fam = d.pop("family")
ndev = torch.cuda.device_count()
for g in [k for k in (1, 2, 4, 8) if k <= ndev]:
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam, nAGQ=th["nAGQ"], device=list(range(g)) if g > 1 else 0)
    eng.find_modes(th["betas"], np.linalg.inv(th["D"]), th["phis"], None, warm=False)
    for _ in range(5):
        r = eng.eval_raw(th["betas"], th["D"], th["phis"], None)
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        r = eng.eval_raw(th["betas"], th["D"], th["phis"], None)
    dt = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    fit = gb.mixed_fit(eng, gb.starting_values_device(eng), {}, hessian=True)
    t_fit = time.perf_counter() - t0
    print(f"devices {g}: {1.0 / dt:8.1f} evaluations/s ({dt * 1e3:.3f} ms each, loglik {r[0]:.6f}); "
          f"mixed_model fit {t_fit:.3f} s (converged {fit['converged']}, betas {np.round(fit['coefficients'], 5)})")
    eng.close()
