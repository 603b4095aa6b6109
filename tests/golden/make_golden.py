#!/usr/bin/env python
"""Freezes the CPU oracle's outputs as committed fixtures.

  python tests/golden/make_golden.py        (from the repo root; CPU only, ~1 min)

Writes
  tests/golden/family_cases.npz   one record per row of tests/family_cases.CASES:
                                  inputs (y, X, Z, id, X_zi, Z_zi, theta), the oracle's grid
                                  (post_modes, packed chol, log_dets) and its outputs (loglik,
                                  score vector of score_mixed, g_beta, g_phi, g_gamma, S,
                                  per-cluster log-likelihood contributions)
  tests/golden/vignette_kat.json  the reference's published known-answer values (rendered
                                  vignettes under docs/articles of the reference, transcribed by
                                  hand with file:line) together with what the oracle computes
                                  for them on the regenerated vignette data.

The reference itself cannot run here (pure R, no R in the image), so the fixtures are the
ORACLE's numbers (oracle/agq.py, oracle/families.py -- a restatement of R/Fit_Funs.R and
R/Functions.R) pinned to the published values where those exist.  The tests then check
(a) on the CPU, that the oracle still reproduces the fixtures (tests/test_golden.py) and
(b) on the GPU, that the CUDA engine matches them through the C ABI (tests/test_gpu_golden.py).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import agq, families as F, vignette_data  # noqa: E402
from family_cases import CASES, build_case, case_id  # noqa: E402
from helpers import oracle_family, oracle_grid, pack_chol, thetas_vec  # noqa: E402

# Published values: paths relative to the reference tree (docs/articles/*.html), SURVEY appendix D
VIGNETTE_KATS = [
    dict(name="D1_nAGQ11", cite="docs/articles/GLMMadaptive.html:332-338", data="binary", family="binomial",
         q_y=1, nAGQ=11, betas=[-2.61816557, -1.32418678, 0.28514773, 0.04015372], D=[[4.26266211]],
         loglik=-374.51972108, tol=5e-7),
    dict(name="D1_nAGQ15", cite="docs/articles/GLMMadaptive.html:332-338", data="binary", family="binomial",
         q_y=1, nAGQ=15, betas=[-2.61679077, -1.32349829, 0.28501148, 0.04020003], D=[[4.26351799]],
         loglik=-374.51983340, tol=5e-7),
    dict(name="D1_nAGQ21", cite="docs/articles/GLMMadaptive.html:332-338", data="binary", family="binomial",
         q_y=1, nAGQ=21, betas=[-2.6168440, -1.3235049, 0.2850151, 0.0401995], D=[[4.2636311]],
         loglik=-374.5198046, tol=5e-7),
    dict(name="D2_q2", cite="docs/articles/Methods.html:189-192,324-330", data="binary", family="binomial",
         q_y=2, nAGQ=11, betas=[-2.059256329, -1.167646870, 0.261720034, 0.001924039],
         D=[[0.47137731, 0.11092919], [0.11092919, 0.06332475]], loglik=-358.8167, tol=5e-4,
         head_ranef=[[-0.3764536, -0.122065606], [0.6675107, 0.289646664], [0.4843611, 0.215634477],
                     [-0.2083419, -0.006681441], [0.2373342, 0.162463265], [-0.3690096, -0.184228112]],
         ranef_tol=1e-5),
]


def family_fixture():
    out = {}
    names = []
    for case in CASES:
        fam, link, q_y, q_zi, phis = case
        cid = case_id(*case)
        names.append(cid)
        fam_name, d, th = build_case(*case)
        ofam = oracle_family(fam_name, link)
        od, gh = oracle_grid(d, th, ofam)
        tv = thetas_vec(th)
        ref = agq.suff_stats(tv, od, ofam, gh)
        rec = dict(y=d["y"], X=d["X"], Z=d["Z"], id=d["id"].astype(np.int64), betas=th["betas"], D=th["D"],
                   nAGQ=np.array(th["nAGQ"]), thetas=tv,
                   post_modes=gh.post_modes, chol=pack_chol(gh.chol_hessians), log_dets=gh.log_dets,
                   loglik=np.array(ref["loglik"]), S=ref["S"], g_beta=ref["g_beta"],
                   score=agq.score_mixed(tv, od, ofam, gh),
                   ll_i=-agq.logLik_mixed(tv, od, ofam, gh, i_contributions=True))
        for k in ("X_zi", "Z_zi"):
            if d.get(k) is not None:
                rec[k] = d[k]
        for k in ("phis", "gammas"):
            if th.get(k) is not None:
                rec[k] = np.asarray(th[k], float)
        for k in ("g_phi", "g_gamma"):
            if k in ref:
                rec[k] = ref[k]
        for k, v in rec.items():
            out[f"{cid}/{k}"] = np.asarray(v)
        print(f"{cid:45s} loglik {float(ref['loglik']):.12f}")
    out["__cases__"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "family_cases.npz"), **out)


def vignette_fixture():
    v = vignette_data.binary_data()
    recs = []
    for k in VIGNETTE_KATS:
        rec = dict(k)
        data = agq.Data(y=v["y"], X=v["X"], Z=v["Z"][:, :k["q_y"]], id=v["id"])
        fam = F.binomial()
        betas, D = np.array(k["betas"]), np.array(k["D"])
        gh = agq.GHfun(np.zeros((data.n, k["q_y"])), data, fam, betas, np.linalg.inv(D), None, None,
                       k["nAGQ"], k["q_y"])
        th = np.concatenate([betas, agq.chol_transf(D)])
        rec["oracle_loglik"] = float(-agq.logLik_mixed(th, data, fam, gh))
        rec["oracle_score"] = [float(s) for s in agq.score_mixed(th, data, fam, gh)]
        rec["oracle_head_modes"] = gh.post_modes[:6].tolist()
        assert abs(rec["oracle_loglik"] - k["loglik"]) < k["tol"], (k["name"], rec["oracle_loglik"])
        recs.append(rec)
        print(f"{k['name']:12s} published {k['loglik']:.8f} oracle {rec['oracle_loglik']:.8f}")
    with open(os.path.join(HERE, "vignette_kat.json"), "w") as f:
        json.dump(dict(source="drizopoulos/GLMMadaptive v0.9-7 rendered vignettes (docs/articles)", kats=recs), f,
                  indent=1)


if __name__ == "__main__":
    family_fixture()
    vignette_fixture()
