"""CUDA engine vs the CPU oracle through the C ABI (libglmma.so via ctypes).

Tolerances (BASELINE.json north_star / SURVEY.md section 8c):
  * eval with shared (modes, chol): relative 1e-10 on the log-likelihood and on every
    score component / sufficient statistic;
  * Newton modes vs the oracle's restatement of the same algorithm: 1e-8 absolute on
    modes, 1e-8 relative on the Cholesky factors, and 1e-9 relative on the
    log-likelihood evaluated on the engine's own grid.
"""
import numpy as np
import pytest

from oracle import agq
from glmmadaptive_b200 import synth
import glmmadaptive_b200 as gb
from helpers import oracle_data, oracle_family, oracle_grid, pack_chol, thetas_vec, relerr

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def small_case(name, n=40, ni=6, ragged=True, seed=11, family=None):
    d, th = synth.make(name, n=n, ni=ni, seed=seed, ragged=ragged, family=family)
    fam_name = d.pop("family")
    return d, th, fam_name


def make_engine(d, th, fam_name, link=None, **kw):
    return gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, link=link, nAGQ=th["nAGQ"],
                     X_zi=d.get("X_zi"), Z_zi=d.get("Z_zi"), offset=d.get("offset"),
                     offset_zi=d.get("offset_zi"), weights=d.get("weights"), **kw)


def check_eval(d, th, fam_name, link=None, rtol=RTOL):
    fam = oracle_family(fam_name, link)
    od, gh = oracle_grid(d, th, fam)
    eng = make_engine(d, th, fam_name, link)
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    tv = thetas_vec(th)
    ref = agq.suff_stats(tv, od, fam, gh)
    got = eng.eval(th["betas"], th["D"], th.get("phis"), th.get("gammas"))
    assert got["n_nonfinite"] == ref["n_nonfinite"]
    assert relerr(got["loglik"], ref["loglik"]) < rtol
    assert relerr(got["sum_w"], ref["sum_w"]) < 1e-14
    scale = max(1.0, np.max(np.abs(ref["g_beta"])))
    assert np.max(np.abs(got["g_beta"] - ref["g_beta"])) < rtol * scale * 10
    assert relerr(got["S"], ref["S"]) < rtol
    if "g_phi" in ref:
        assert np.max(np.abs(got["g_phi"] - ref["g_phi"])) < rtol * max(1.0, abs(ref["g_phi"][0])) * 10
    if "g_gamma" in ref:
        sc = max(1.0, np.max(np.abs(ref["g_gamma"])))
        assert np.max(np.abs(got["g_gamma"] - ref["g_gamma"])) < rtol * sc * 10
    # the reference-shaped closures
    nll = gb.logLik_mixed(tv, eng, None)
    assert relerr(nll, agq.logLik_mixed(tv, od, fam, gh)) < rtol
    sc_ref = agq.score_mixed(tv, od, fam, gh)
    sc_got = gb.score_mixed(tv, eng, None)
    assert np.max(np.abs(sc_got - sc_ref)) < rtol * 10 * max(1.0, np.max(np.abs(sc_ref)))
    # loglik-only kernel agrees with the fused one
    ll_only = eng.eval(th["betas"], th["D"], th.get("phis"), th.get("gammas"), want_score=False)
    assert relerr(ll_only["loglik"], got["loglik"]) < 1e-13
    eng.close()
    return got, ref


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C5a", "C5b"])
def test_eval_parity_core_families(name):
    """Every BASELINE.json configuration's own shape: family, D, phis, nAGQ (11; 7 for the q = 3
    zero-inflated model), censoring pattern -- at a size the oracle finishes in seconds."""
    d, th, fam_name = small_case(name)
    check_eval(d, th, fam_name)


@pytest.mark.parametrize("name", ["C4", "C5a", "C5b"])
def test_eval_parity_baseline_configs_complete_clusters(name):
    """The same configurations with the benchmark's complete clusters of 8 observations (the register
    variant's one-block-of-8 path) and the engine's own Newton modes on the oracle's grid."""
    d, th, fam_name = small_case(name, n=48, ni=8, ragged=False, seed=5)
    check_eval(d, th, fam_name)
    fam = oracle_family(fam_name)
    od, gh = oracle_grid(d, th, fam, mode_finder="newton")
    eng = make_engine(d, th, fam_name)
    st = eng.find_modes(th["betas"], np.linalg.inv(th["D"]), th.get("phis"), th.get("gammas"), warm=False)
    assert st["n_not_converged"] == 0 and st["n_chol_failed"] == 0
    modes, chol, logdet = eng.get_modes()
    assert np.max(np.abs(modes - gh.post_modes)) < 1e-7
    assert np.max(np.abs(logdet - gh.log_dets)) < 1e-7
    eng.close()


def test_eval_parity_weights_offset():
    d, th, fam_name = small_case("C3", n=30)
    rng = np.random.default_rng(5)
    d["weights"] = rng.uniform(0.5, 2.0, size=30)
    d["offset"] = rng.normal(0, 0.3, size=len(d["y"]))
    check_eval(d, th, fam_name)


def test_eval_large_cluster_streaming_path():
    # clusters larger than the shared-memory observation tile take the streaming path
    d, th, fam_name = small_case("C3", n=12, ni=70, ragged=True, seed=3)
    check_eval(d, th, fam_name)


def test_modes_vs_oracle_newton():
    for name in ["C1", "C2", "C3"]:
        d, th, fam_name = small_case(name, n=30)
        fam = oracle_family(fam_name)
        od, gh = oracle_grid(d, th, fam, mode_finder="newton")
        eng = make_engine(d, th, fam_name)
        invD = np.linalg.inv(th["D"])
        st = eng.find_modes(th["betas"], invD, th.get("phis"), th.get("gammas"), warm=False)
        assert st["n_not_converged"] == 0 and st["n_chol_failed"] == 0
        modes, chol, logdet = eng.get_modes()
        assert np.max(np.abs(modes - gh.post_modes)) < 1e-8
        assert relerr(chol, pack_chol(gh.chol_hessians)) < 1e-7
        assert np.max(np.abs(logdet - gh.log_dets)) < 1e-8
        tv = thetas_vec(th)
        got = eng.eval(th["betas"], th["D"], th.get("phis"), th.get("gammas"))
        assert relerr(got["loglik"], -agq.logLik_mixed(tv, od, fam, gh)) < 1e-9
        eng.close()


def test_ghfun_mirror_and_bfgs_reference_path():
    # end to end through GHfun: engine (Newton) vs the oracle's reference-faithful
    # BFGS + cd_vec path (R/Functions.R:80,90): documented looser bound 1e-8 on logLik
    d, th, fam_name = small_case("C1", n=40)
    fam = oracle_family(fam_name)
    od, gh_ref = oracle_grid(d, th, fam, mode_finder="bfgs")
    eng = make_engine(d, th, fam_name)
    gh = gb.GHfun(eng, np.zeros((eng.n, eng.q)), th["betas"], np.linalg.inv(th["D"]))
    tv = thetas_vec(th)
    assert relerr(gb.logLik_mixed(tv, eng, gh), agq.logLik_mixed(tv, od, fam, gh_ref)) < 1e-8
    assert np.max(np.abs(gh.post_modes - gh_ref.post_modes)) < 1e-3
    pv = gh.post_vars
    assert relerr(pv[0], gh_ref.post_vars[0]) < 1e-3
    eng.close()
