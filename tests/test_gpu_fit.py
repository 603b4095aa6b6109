"""GPU: the EM-phase entry points and full mixed_model() fits of the engine against the oracle's
restatement of R/mixed_fit.R, and against the coefficients the reference's vignette publishes.

Tolerances: E-step / M-step scores with shared grid: relative 1e-10 (as eval); fitted
coefficients: 1e-6 (north_star) against the oracle's fit with the same mode finder, and every
printed digit of the published `fm1` estimates."""
import json
import os

import numpy as np
import pytest

from oracle import agq, families as F, vignette_data, fit as ofit
import glmmadaptive_b200 as gb
from glmmadaptive_b200 import synth
from helpers import oracle_data, oracle_family, oracle_grid, pack_chol, thetas_vec, relerr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vignette_kat.json")


def _case(name, n, ni=6, **kw):
    d, th = synth.make(name, n=n, ni=ni, seed=21, ragged=True, **kw)
    fam_name = d.pop("family")
    return d, th, fam_name


@pytest.mark.parametrize("name", ["C1", "C3", "C4"])
def test_em_estep_and_frozen_scores(name):
    d, th, fam_name = _case(name, 40)
    if name == "C3":
        d["weights"] = np.random.default_rng(1).uniform(0.5, 2.0, 40)
    _em_check(d, th, fam_name)


@pytest.mark.parametrize("name", ["C3", "C4"])
def test_em_passes_on_clusters_larger_than_the_staged_tile(name):
    """Cluster sizes 1 + Poisson(44): nearly all beyond the generic kernel's 32 staged observations
    (tile-outer path: node sums per warp, two passes over the tiles), E-step and frozen-weight scores."""
    d, th, fam_name = _case(name, 14, ni=45)
    sizes = np.bincount(np.unique(d["id"], return_inverse=True)[1])
    assert (sizes > 32).sum() >= 10
    _em_check(d, th, fam_name)


def _em_check(d, th, fam_name):
    fam = oracle_family(fam_name)
    od, gh = oracle_grid(d, th, fam)
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, nAGQ=th["nAGQ"], X_zi=d.get("X_zi"),
                    Z_zi=d.get("Z_zi"), weights=d.get("weights"))
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    betas, D, phis, gammas = th["betas"], th["D"], th.get("phis"), th.get("gammas")
    p_by, post_b2, ll, eta_zi = ofit.e_step(od, fam, gh, betas, D, phis, gammas)
    es = eng.em_estep(betas, D, phis, gammas)
    assert relerr(es["loglik"], ll) < 1e-10
    assert relerr(es["S"].reshape(-1), np.sum(post_b2, axis=0)) < 1e-10
    # the E-step result is the fused evaluation's
    ev = eng.eval(betas, D, phis, gammas)
    assert relerr(es["loglik"], ev["loglik"]) < 1e-13 and relerr(es["g_beta"], ev["g_beta"]) < 1e-9
    # scores at a DIFFERENT theta with the posterior weights frozen at the E-step's theta
    b2 = betas + np.array([0.05, -0.03, 0.02])
    ph2 = None if phis is None else phis + 0.07
    g2 = None if gammas is None else gammas + np.array([0.04, -0.06])
    sc = eng.em_score(b2, ph2, g2)
    ref_b = ofit.score_betas(b2, od, fam, gh, ph2, eta_zi if g2 is None else ofit._etas(od, gh, b2, g2)[1], p_by)
    assert np.max(np.abs(-sc["g_beta"] - ref_b)) < 1e-10 * max(1.0, np.max(np.abs(ref_b))) * 10
    if phis is not None:
        ref_p = ofit.score_phis(ph2, od, fam, gh, b2, eta_zi, p_by)
        if d.get("weights") is None:       # the reference drops the weights here (Fit_Funs.R:392-394)
            assert abs(-sc["g_phi"][0] - ref_p[0]) < 1e-9 * max(1.0, abs(ref_p[0]))
    if gammas is not None:
        ref_g = ofit.score_gammas(g2, od, fam, gh, b2, ph2, p_by)
        assert np.max(np.abs(-sc["g_gamma"] - ref_g)) < 1e-9 * max(1.0, np.max(np.abs(ref_g)))
    eng.close()


@pytest.mark.parametrize("name", ["D1_nAGQ11", "D1_nAGQ15"])
def test_mixed_model_reproduces_published_fm1(name):
    """mixed_model(y ~ sex * time, random = ~ 1 | id, family = binomial()) on the vignette data:
    the engine's fit prints the same digits as the reference's rendered vignette."""
    k = [r for r in json.load(open(GOLD))["kats"] if r["name"] == name][0]
    v = vignette_data.binary_data()
    fit = gb.mixed_model(v["y"], v["X"], v["Z"][:, :1], v["id"], "binomial", control=dict(nAGQ=k["nAGQ"]))
    assert fit["converged"]
    assert np.max(np.abs(fit["coefficients"] - np.array(k["betas"]))) < 1e-6
    assert abs(fit["D"][0, 0] - k["D"][0][0]) < 1e-6
    assert abs(fit["logLik"] - k["loglik"]) < 5e-7
    # standard errors from the engine's Hessian: sqrt(diag(solve(Hessian))) is finite and positive
    se = np.sqrt(np.diag(np.linalg.inv(fit["Hessian"])))
    assert np.all(np.isfinite(se)) and np.all(se > 0)


@pytest.mark.parametrize("name,n", [("C3", 60), ("C2", 50), ("C4", 60)])
def test_mixed_model_vs_oracle_fit(name, n):
    d, th, fam_name = _case(name, n)
    fam = oracle_family(fam_name)
    od = oracle_data(d)
    con = dict(nAGQ=7 if name == "C4" else 9, iter_EM=12, iter_qN_outer=4)
    ref = ofit.mixed_fit(od, fam, control=con, mode_finder="newton", hessian=False)
    fit = gb.mixed_model(d["y"], d["X"], d["Z"], d["id"], fam_name, X_zi=d.get("X_zi"), Z_zi=d.get("Z_zi"),
                         control=con, hessian=False)
    assert fit["converged"] == ref["converged"]
    assert np.max(np.abs(fit["coefficients"] - ref["coefficients"])) < 1e-6
    assert np.max(np.abs(fit["D"] - ref["D"])) < 1e-6
    if ref["phis"] is not None:
        assert np.max(np.abs(fit["phis"] - ref["phis"])) < 1e-6
    if ref["gammas"] is not None:
        assert np.max(np.abs(fit["gammas"] - ref["gammas"])) < 1e-6
    assert abs(fit["logLik"] - ref["logLik"]) < 1e-7 * abs(ref["logLik"])


def test_device_glm_starts_match_host_glm_fit():
    """glm.fit starting values computed by IRLS passes on the device (glmma_glm_pass) against the
    host restatement of stats::glm.fit, for every start mixed_model() computes
    (R/mixed_model.R:157-194)."""
    from glmmadaptive_b200 import fit as hf
    from family_cases import build_case
    for case in [("binomial", "logit", 2, 0, None), ("binomial2", "logit", 2, 0, None), ("poisson", None, 1, 0, None),
                 ("negative.binomial", None, 1, 0, [0.4]), ("Gamma.fam", None, 2, 0, [1.0]),
                 ("zi.poisson", None, 2, 1, None), ("hurdle.negative.binomial", None, 2, 1, [0.6]),
                 ("hurdle.lognormal", None, 1, 1, [-0.3]), ("hurdle.beta.fam", None, 1, 1, [1.5]),
                 ("zi.binomial", None, 1, 1, None)]:
        fam_name, d, th = build_case(*case, n=200)
        rng = np.random.default_rng(5)
        off = rng.normal(0, 0.2, len(d["X"]))
        eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, link=case[1], nAGQ=5, X_zi=d.get("X_zi"),
                        Z_zi=d.get("Z_zi"), offset=off, offset_zi=None if d.get("X_zi") is None else 0.5 * off)
        dev = hf.starting_values_device(eng)
        host = hf.starting_values(d["y"], d["X"], d.get("X_zi"), eng.spec.family, eng.q, eng.n_phis, False,
                                  off, None if d.get("X_zi") is None else 0.5 * off)
        for k in host:
            assert np.max(np.abs(np.asarray(dev[k]) - np.asarray(host[k]))) < 1e-8, (case[0], k)
        eng.close()


def test_predict_subject_specific_vs_oracle():
    """predict(type = "subject_specific") point predictions for new clusters: EB modes from the
    device mode search, then the reference's formulas (R/methods.R:976-1012)."""
    for name in ("C3", "C4"):
        d, th, fam_name = _case(name, 30)
        fam = oracle_family(fam_name)
        od, gh = oracle_grid(d, th, fam, mode_finder="newton")
        fit = dict(coefficients=th["betas"], D=th["D"], phis=th.get("phis"), gammas=th.get("gammas"))
        pr = gb.predict_subject_specific(fit, d["y"], d["X"], d["Z"], d["id"], fam_name, X_zi=d.get("X_zi"),
                                         Z_zi=d.get("Z_zi"), nAGQ=th["nAGQ"])
        assert np.max(np.abs(pr["post_modes"] - gh.post_modes)) < 1e-7
        idx = np.repeat(np.arange(od.n), np.diff(od.row_ptr))
        eta = od.X @ th["betas"] + np.sum(od.Z * gh.post_modes[idx, :od.q_y], axis=1)
        ref = np.exp(eta)
        if name == "C4":
            ez = od.X_zi @ th["gammas"] + od.Z_zi[:, 0] * gh.post_modes[idx, 2]
            ref = ref / (1 + np.exp(ez))
        assert np.max(np.abs(pr["pred"] - ref) / ref) < 1e-6


@pytest.mark.parametrize("name", ["km1", "km2", "hm1", "hm2", "dm2"])
def test_mixed_model_reproduces_published_two_part_fits(name):
    """The engine's mixed_model() on the two-part vignette data prints the vignette's estimates
    (hurdle log-normal, hurdle negative binomial, hurdle poisson; q up to 3)."""
    from kat_hurdle import FIT_KATS, kat_inputs, check_fit_against_kat
    k = [r for r in FIT_KATS if r["name"] == name][0]
    d = kat_inputs(k)
    fit = gb.mixed_model(d["y"], d["X"], d["Z"], d["id"], k["family"], X_zi=d["X_zi"], Z_zi=d["Z_zi"],
                         control=dict(nAGQ=k["nAGQ"]), diag_D=k["diag_D"], hessian=False)
    check_fit_against_kat(k, fit)


def test_predictions_match_published_vignette_values():
    """docs/articles/Methods.html of the reference: head(fitted(fm, type = "subject_specific")) (:389-394)
    and predict(fm, newdata = pred_DF, type = "subject_specific") for subject 1's first four visits as a
    new patient (:546-551), at the published estimates of fm (binomial, random intercepts and slopes).
    The fitted values agree to every printed digit; the prediction to 5e-6 (the reference finds the new
    patient's mode with optim's default BFGS tolerance, the engine with Newton to 1e-10)."""
    v = vignette_data.binary_data()
    fm = dict(coefficients=np.array([-2.059256329, -1.167646870, 0.261720034, 0.001924039]),
              D=np.array([[0.47137731, 0.11092919], [0.11092919, 0.06332475]]), phis=None, gammas=None)
    pr = gb.predict_subject_specific(fm, v["y"], v["X"], v["Z"], v["id"], "binomial", nAGQ=11)
    pub = np.array([0.08048986, 0.08197442, 0.09997322, 0.23877793, 0.24377249, 0.24418982])
    assert np.max(np.abs(pr["pred"][:6] - pub)) < 2e-8
    new = gb.predict_subject_specific(fm, v["y"][:4], v["X"][:4], v["Z"][:4], np.zeros(4, int), "binomial", nAGQ=11)
    assert np.max(np.abs(new["pred"] - np.array([0.07876144, 0.07981214, 0.09220958, 0.17714850]))) < 5e-6


def test_mixed_model_reproduces_published_standard_errors():
    """summary(fm) of the Multiple_Comparisons vignette: the engine's estimates, the standard errors from
    its Hessian (cd_vec of score_mixed, 2 * nparams fused evaluations), sd and logLik."""
    from test_fit_oracle import check_multcomp
    v = vignette_data.multcomp_binary_data()
    check_multcomp(gb.mixed_model(v["y"], v["X"], v["Z"], v["id"], "binomial", hessian=True))


def _family_kats():
    from kat_families import FAMILY_KATS
    return FAMILY_KATS


@pytest.mark.parametrize("name", [k["name"] for k in _family_kats()])
def test_engine_reproduces_published_family_fits(name):
    """Families of BASELINE configs C2-C5 against real R output (tests/kat_families.py): the engine's
    log-likelihood at the published estimates prints the published value, and the engine's whole
    mixed_model() fit lands on the published estimates (poisson, zi.poisson q = 1 / 2,
    zi.negative.binomial, negative binomial, beta)."""
    from kat_families import kat_data, kat_theta, check_family_fit
    k = [r for r in _family_kats() if r["name"] == name][0]
    d = kat_data(k)
    betas, D, phis, gammas = kat_theta(k)
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], k["family"], nAGQ=k["nAGQ"], X_zi=d.get("X_zi"),
                    Z_zi=d.get("Z_zi"))
    st = eng.find_modes(betas, np.linalg.inv(D), phis, gammas, warm=False)
    assert st["n_not_converged"] == 0 and st["n_chol_failed"] == 0
    got = eng.eval(betas, D, phis, gammas)
    assert abs(got["loglik"] - k["loglik"]) < k["eval_tol"], (name, got["loglik"])
    eng.close()
    fit = gb.mixed_model(d["y"], d["X"], d["Z"], d["id"], k["family"], X_zi=d.get("X_zi"), Z_zi=d.get("Z_zi"),
                         control=dict(nAGQ=k["nAGQ"]), hessian=False)
    check_family_fit(k, fit, newton=True)


@pytest.mark.parametrize("name", ["C1", "C3", "C4", "C5a"])
def test_log_post_b_batch_vs_oracle(name):
    """glmma_log_post_b: log_post_b of predict()'s Metropolis step (R/methods.R:1020-1036) for one
    proposal per cluster, against the oracle's restatement of the closure (relative 1e-10)."""
    d, th, fam_name = _case(name, 35)
    fam = oracle_family(fam_name)
    od = oracle_data(d)
    rng = np.random.default_rng(3)
    b = rng.normal(0, 0.6, size=(od.n, od.nRE))
    invD = np.linalg.inv(th["D"])
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, nAGQ=th["nAGQ"], X_zi=d.get("X_zi"), Z_zi=d.get("Z_zi"))
    got = eng.log_post_b(b, th["betas"], invD, th.get("phis"), th.get("gammas"))
    ref = np.array([-agq.make_log_post(od, fam, i, th["betas"], invD, th.get("phis"), th.get("gammas"))[0](b[i])
                    for i in range(od.n)])
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) < 1e-10
    eng.close()


@pytest.mark.parametrize("name,diag", [("C1", False), ("C2", False), ("C3", False), ("C3", True), ("C4", False),
                                        ("C5a", False), ("C5b", False)])
def test_observed_information_vs_cd_vec_of_the_score(name, diag):
    """glmma_observed_information (one pass: E[d2a] + Cov(da) per cluster) against what the reference
    computes, cd_vec of score_mixed on the same grid (R/mixed_fit.R:324-342) -- through the engine's score
    and through the oracle's -- to 1e-5 of the largest entry (cd_vec's own truncation error is ~1e-6)."""
    from glmmadaptive_b200.fit import cd_vec
    d, th, fam_name = _case(name, 50)
    if name == "C3" and not diag:
        d["weights"] = np.random.default_rng(4).uniform(0.5, 2.0, 50)
    fam = oracle_family(fam_name)
    if diag:
        th = dict(th); th["D"] = np.diag(np.diag(th["D"]))
    od, gh = oracle_grid(d, th, fam)
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, nAGQ=th["nAGQ"], X_zi=d.get("X_zi"),
                    Z_zi=d.get("Z_zi"), weights=d.get("weights"))
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    tv = thetas_vec(th, diag_D=diag)
    H = eng.observed_information(th["betas"], th["D"], th.get("phis"), th.get("gammas"), diag_D=diag)
    assert np.allclose(H, H.T)
    ref = cd_vec(tv, lambda t: gb.score_mixed(t, eng, None, diag))
    ref_o = agq.cd_vec(tv, lambda t: agq.score_mixed(t, od, fam, gh, diag_D=diag))
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(ref - ref_o)) < 1e-7 * scale
    # entry-wise against the scale of its row and column (blocks differ by orders of magnitude)
    sd = np.sqrt(np.abs(np.diag(ref)))
    assert np.max(np.abs(H - ref) / np.outer(sd, sd)) < 2e-5, np.max(np.abs(H - ref) / np.outer(sd, sd))
    eng.close()
