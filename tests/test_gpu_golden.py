"""GPU: the CUDA engine, through the C ABI, against the committed golden fixtures
(tests/golden/family_cases.npz: inputs, grid and outputs frozen from the oracle;
tests/golden/vignette_kat.json: values published by the reference's vignettes).
Nothing under oracle/ is needed at run time except the vignette data generator.
Tolerance: relative 1e-10 on logLik / score / S (north_star), 5e-7 .. 5e-4 on the KATs
(the digits the vignettes print)."""
import json
import os

import numpy as np
import pytest

import glmmadaptive_b200 as gb
from family_cases import CASES, case_id
from test_golden import GOLD, load_case

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope="module")
def npz():
    return np.load(os.path.join(GOLD, "family_cases.npz"))


@pytest.mark.parametrize("case", CASES, ids=lambda c: case_id(*c))
def test_engine_matches_fixture(npz, case):
    fam, link, q_y, q_zi, phis = case
    g = load_case(npz, case_id(*case))
    fam_name = {"binomial2": "binomial", "beta.binomial2": "beta.binomial"}.get(fam, fam)
    eng = gb.Engine(g["y"], g["X"], g["Z"], g["id"], fam_name, link=link, nAGQ=int(g["nAGQ"]),
                    X_zi=g.get("X_zi"), Z_zi=g.get("Z_zi"), df=4.0 if fam == "students.t" else None)
    eng.set_modes(g["post_modes"], g["chol"])
    got = eng.eval(g["betas"], g["D"], g.get("phis"), g.get("gammas"))
    assert got["n_nonfinite"] == 0
    assert abs(got["loglik"] - float(g["loglik"])) < RTOL * abs(float(g["loglik"]))
    for key in ("g_beta", "g_phi", "g_gamma"):
        if key in g:
            assert np.max(np.abs(got[key] - g[key])) < 10 * RTOL * max(1.0, np.max(np.abs(g[key]))), key
    assert np.max(np.abs(got["S"] - g["S"]) / np.abs(g["S"])) < RTOL
    sc = gb.score_mixed(g["thetas"], eng, None)
    assert np.max(np.abs(sc - g["score"])) < 10 * RTOL * max(1.0, np.max(np.abs(g["score"])))
    ll_i = -gb.logLik_mixed(g["thetas"], eng, None, i_contributions=True)
    assert np.max(np.abs(ll_i - g["ll_i"]) / np.maximum(np.abs(g["ll_i"]), 1.0)) < RTOL
    # the engine's own mode search lands on the fixture's grid
    st = eng.find_modes(g["betas"], np.linalg.inv(g["D"]), g.get("phis"), g.get("gammas"), warm=False)
    assert st["n_chol_failed"] == 0 and st["n_not_converged"] == 0
    modes, chol, logdet = eng.get_modes()
    assert np.max(np.abs(modes - g["post_modes"])) < 1e-7
    assert np.max(np.abs(logdet - g["log_dets"])) < 1e-7
    eng.close()


def test_engine_matches_published_vignette_values():
    from oracle import vignette_data          # data generator only (R RNG restatement)
    v = vignette_data.binary_data()
    for k in json.load(open(os.path.join(GOLD, "vignette_kat.json")))["kats"]:
        q = k["q_y"]
        eng = gb.Engine(v["y"], v["X"], v["Z"][:, :q], v["id"], k["family"], nAGQ=k["nAGQ"])
        betas, D = np.array(k["betas"]), np.array(k["D"])
        gh = gb.GHfun(eng, np.zeros((eng.n, q)), betas, np.linalg.inv(D))
        tv = np.concatenate([betas, gb.chol_transf(D)])
        ll = -gb.logLik_mixed(tv, eng, gh)
        assert abs(ll - k["loglik"]) < k["tol"], (k["name"], ll)
        assert abs(ll - k["oracle_loglik"]) < 1e-8 * abs(ll)
        if "head_ranef" in k:
            assert np.max(np.abs(gh.post_modes[:6] - np.array(k["head_ranef"]))) < k["ranef_tol"]
        eng.close()
