"""Every accelerated family through the C ABI against the oracle (shared modes/chol):
relative 1e-10 on the log-likelihood and on each score component (north_star tolerance).
Also covers random effects in the zero part (q_zi = 1), two-column binomial responses,
probit/cloglog links, the vignette hurdle data (KAT D4/D5 shapes) and ragged clusters."""
import numpy as np
import pytest

from oracle import agq, vignette_data
import glmmadaptive_b200 as gb
from helpers import oracle_data, oracle_family, oracle_grid, pack_chol, thetas_vec, relerr

pytestmark = pytest.mark.gpu
RTOL = 1e-10


from family_cases import CASES, build_case, case_id, design, response


@pytest.mark.parametrize("fam,link,q_y,q_zi,phis", CASES)
def test_family_parity(fam, link, q_y, q_zi, phis):
    fam_name, d, th = build_case(fam, link, q_y, q_zi, phis)
    ofam = oracle_family(fam_name, link)
    D = th["D"]
    od, gh = oracle_grid(d, th, ofam)
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, link=link, nAGQ=th["nAGQ"],
                    X_zi=d.get("X_zi"), Z_zi=d.get("Z_zi"), df=th.get("df"))
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    tv = thetas_vec(th)
    ref = agq.suff_stats(tv, od, ofam, gh)
    got = eng.eval(th["betas"], th["D"], th["phis"], th["gammas"])
    assert got["n_nonfinite"] == ref["n_nonfinite"] == 0
    assert relerr(got["loglik"], ref["loglik"]) < RTOL
    for key in ("g_beta", "g_phi", "g_gamma"):
        if key in ref:
            sc = max(1.0, np.max(np.abs(ref[key])))
            assert np.max(np.abs(got[key] - ref[key])) < 10 * RTOL * sc, key
    assert relerr(got["S"], ref["S"]) < RTOL
    sc_ref = agq.score_mixed(tv, od, ofam, gh)
    sc_got = gb.score_mixed(tv, eng, None)
    assert np.max(np.abs(sc_got - sc_ref)) < 10 * RTOL * max(1.0, np.max(np.abs(sc_ref)))
    # mode search on the engine agrees with the oracle's restatement of the same algorithm
    st = eng.find_modes(th["betas"], np.linalg.inv(D), th["phis"], th["gammas"], warm=False)
    assert st["n_chol_failed"] == 0
    modes, chol, logdet = eng.get_modes()
    assert np.max(np.abs(modes - gh.post_modes)) < 1e-7
    assert np.max(np.abs(logdet - gh.log_dets)) < 1e-7
    eng.close()


def test_diag_D_and_contributions():
    rng = np.random.default_rng(3)
    ofam = oracle_family("negative.binomial")
    d, N = design(rng, 30, 5, 2)
    D = np.diag([0.6, 0.3])
    betas = np.array([0.3, 0.5, -0.4])
    d["y"] = response("negative.binomial", rng, N, d["X"] @ betas)
    d["weights"] = rng.uniform(0.5, 1.5, 30)
    th = dict(betas=betas, D=D, phis=np.array([0.3]), gammas=None, nAGQ=11)
    od, gh = oracle_grid(d, th, ofam)
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], "negative.binomial", nAGQ=11, weights=d["weights"])
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    tv = thetas_vec(th, diag_D=True)
    assert relerr(gb.logLik_mixed(tv, eng, None, diag_D=True), agq.logLik_mixed(tv, od, ofam, gh, diag_D=True)) < RTOL
    ref = agq.score_mixed(tv, od, ofam, gh, diag_D=True)
    got = gb.score_mixed(tv, eng, None, diag_D=True)
    assert np.max(np.abs(got - ref)) < 10 * RTOL * max(1.0, np.max(np.abs(ref)))
    # per-cluster / per-observation contributions (i_contributions = TRUE)
    tvf = thetas_vec(th)
    ll_i = gb.logLik_mixed(tvf, eng, None, i_contributions=True)
    assert relerr(ll_i, agq.logLik_mixed(tvf, od, ofam, gh, i_contributions=True)) < RTOL
    cr = agq.score_mixed(tvf, od, ofam, gh, i_contributions=True)
    cg = gb.score_mixed(tvf, eng, None, i_contributions=True)
    w_obs = d["weights"][d["id"]]
    assert np.max(np.abs(cg["score_betas"] - cr["score_betas"] * w_obs[:, None])) < 1e-9
    assert np.max(np.abs(cg["score_phis"] - cr["score_phis"] * w_obs)) < 1e-9
    eng.close()


def test_kat_vignette_binomial_published_loglik():
    """KAT D1/D2: the published logLik of the reference's vignette fits,
    evaluated by the ENGINE end to end (its own mode search) at the published estimates."""
    v = vignette_data.binary_data()
    # D1: ~ 1 | id, nAGQ = 11 / 15 / 21  (docs/articles/GLMMadaptive.html:332-338)
    for nagq, betas, Dv, pub in [
        (11, [-2.61816557, -1.32418678, 0.28514773, 0.04015372], 4.26266211, -374.51972108),
        (15, [-2.61679077, -1.32349829, 0.28501148, 0.04020003], 4.26351799, -374.51983340),
        (21, [-2.6168440, -1.3235049, 0.2850151, 0.0401995], 4.2636311, -374.5198046),
    ]:
        eng = gb.Engine(v["y"], v["X"], v["Z"][:, :1], v["id"], "binomial", nAGQ=nagq)
        D = np.array([[Dv]])
        gh = gb.GHfun(eng, np.zeros((eng.n, 1)), np.array(betas), np.linalg.inv(D))
        tv = np.concatenate([betas, gb.chol_transf(D)])
        assert abs(-gb.logLik_mixed(tv, eng, gh) - pub) < 5e-7
        eng.close()
    # D2: ~ time | id, q = 2 (docs/articles/Methods.html:189-192, 324-330)
    betas = np.array([-2.059256329, -1.167646870, 0.261720034, 0.001924039])
    D = np.array([[0.47137731, 0.11092919], [0.11092919, 0.06332475]])
    eng = gb.Engine(v["y"], v["X"], v["Z"], v["id"], "binomial", nAGQ=11)
    gh = gb.GHfun(eng, np.zeros((eng.n, 2)), betas, np.linalg.inv(D))
    tv = np.concatenate([betas, gb.chol_transf(D)])
    assert abs(-gb.logLik_mixed(tv, eng, gh) - (-358.8167)) < 5e-4
    pub = np.array([[-0.3764536, -0.122065606], [0.6675107, 0.289646664], [0.4843611, 0.215634477],
                    [-0.2083419, -0.006681441], [0.2373342, 0.162463265], [-0.3690096, -0.184228112]])
    assert np.max(np.abs(gh.post_modes[:6] - pub)) < 1e-5      # head(ranef(fm)), Methods.html:324-330
    eng.close()


def test_kat_hurdle_poisson_q3():
    """KAT D5: hurdle poisson, q = 3 (random effect in the zero part), nAGQ = 7, K = 343
    (docs/articles/ZeroInflated_and_TwoPart_Models.html:682-706): published logLik -2797.307."""
    v = vignette_data.hurdle_count_data()
    betas = np.array([1.52792269, 0.08389745, 0.09206185, -0.08020395])
    gammas = np.array([-1.5398382, 0.5175922])
    sd = np.array([0.8226, 0.3644, 0.6584])
    R = np.array([[1, -0.3125, -0.0065], [-0.3125, 1, 0.0636], [-0.0065, 0.0636, 1]])
    D = R * np.outer(sd, sd)
    eng = gb.Engine(v["y"], v["X"], v["Z"], v["id"], "hurdle.poisson", nAGQ=7, X_zi=v["X_zi"], Z_zi=v["Z_zi"])
    gh = gb.GHfun(eng, np.zeros((eng.n, 3)), betas, np.linalg.inv(D), None, gammas)
    assert gh.stats["n_not_converged"] == 0
    tv = np.concatenate([betas, gb.chol_transf(D), gammas])
    # sds/correlations are published with 4 digits only
    assert abs(-gb.logLik_mixed(tv, eng, gh) - (-2797.307)) < 5e-2
    eng.close()


def test_censored_normal_floor_regimes():
    """censored.normal with censoring points far in the tails (|t| up to ~9): the sqrt(eps) floors of
    the reference's score functions (R/Fit_Funs.R:1395, 1413, 1421) become active; the register
    variant's table-based Gaussian tail functions must reproduce them (and the plain regime) to 1e-10."""
    rng = np.random.default_rng(17)
    ofam = oracle_family("censored.normal")
    d, N = design(rng, 40, 6, 2)
    betas = np.array([0.3, 0.5, -0.4])
    D = np.array([[0.3, 0.05], [0.05, 0.2]])
    eta = d["X"] @ betas
    sigma = 0.7
    v = eta + sigma * rng.standard_normal(N)
    ind = rng.integers(0, 3, N).astype(float)
    # push a third of the censored observations deep into either tail
    shift = np.where(rng.random(N) < 0.35, rng.choice([-1.0, 1.0], N) * rng.uniform(3.5, 6.0, N), 0.0)
    y = np.column_stack([v + np.where(ind > 0, shift, 0.0), ind])
    d["y"] = y
    th = dict(betas=betas, D=D, phis=np.array([np.log(sigma)]), gammas=None, nAGQ=11)
    od, gh = oracle_grid(d, th, ofam)
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], "censored.normal", nAGQ=11)
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    tv = thetas_vec(th)
    ref = agq.suff_stats(tv, od, ofam, gh)
    got = eng.eval(th["betas"], th["D"], th["phis"], None)
    assert got["n_nonfinite"] == ref["n_nonfinite"] == 0
    assert relerr(got["loglik"], ref["loglik"]) < RTOL
    for key in ("g_beta", "g_phi"):
        sc = max(1.0, np.max(np.abs(ref[key])))
        assert np.max(np.abs(got[key] - ref[key])) < 10 * RTOL * sc, key
    assert relerr(got["S"], ref["S"]) < RTOL
    # some grid points really are in the floored regime
    tmin = np.min(np.abs((y[ind > 0, 0][:, None] - (od.X[ind > 0] @ betas)[:, None] - gh.Ztb[ind > 0]) / sigma))
    tmax = np.max(np.abs((y[ind > 0, 0][:, None] - (od.X[ind > 0] @ betas)[:, None] - gh.Ztb[ind > 0]) / sigma))
    assert tmax > 6.0
    eng.close()


def test_contributions_zero_inflated_q3_diag():
    """i_contributions = TRUE pieces (per-cluster logLik, per-observation beta / gamma score rows,
    per-cluster D rows) and the diagonal-D parameterisation for a zero-inflated family with a random
    effect in the zero part (q = 3)."""
    fam_name, d, th = build_case("zi.poisson", None, 2, 1, None, n=30)
    ofam = oracle_family(fam_name)
    th["D"] = np.diag(np.diag(th["D"]))
    od, gh = oracle_grid(d, th, ofam)
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, nAGQ=th["nAGQ"], X_zi=d["X_zi"], Z_zi=d["Z_zi"])
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    for diag in (True, False):
        tv = thetas_vec(th, diag_D=diag)
        assert relerr(gb.logLik_mixed(tv, eng, None, diag_D=diag), agq.logLik_mixed(tv, od, ofam, gh, diag_D=diag)) < RTOL
        ref = agq.score_mixed(tv, od, ofam, gh, diag_D=diag)
        got = gb.score_mixed(tv, eng, None, diag_D=diag)
        assert np.max(np.abs(got - ref)) < 10 * RTOL * max(1.0, np.max(np.abs(ref)))
    tv = thetas_vec(th)
    ll_i = gb.logLik_mixed(tv, eng, None, i_contributions=True)
    assert relerr(ll_i, agq.logLik_mixed(tv, od, ofam, gh, i_contributions=True)) < RTOL
    cr = agq.score_mixed(tv, od, ofam, gh, i_contributions=True)
    cg = gb.score_mixed(tv, eng, None, i_contributions=True)
    assert np.max(np.abs(cg["score_betas"] - cr["score_betas"])) < 1e-9
    assert np.max(np.abs(cg["score_gammas"] - cr["score_gammas"])) < 1e-9
    assert np.max(np.abs(cg["score_D"] - cr["score_D"])) < 1e-9
    # the rows add up to the totals
    tot = gb.score_mixed(tv, eng, None)
    assert np.max(np.abs(cg["score_betas"].sum(axis=0) - tot[:3])) < 1e-9
    eng.close()


def test_kat_hurdle_fits_published_loglik():
    """KATs km2 / hm1 / hm2 (tests/kat_hurdle.py): the ENGINE, with its own mode search, at the
    published estimates of the reference's two-part vignette."""
    from kat_hurdle import HURDLE_KATS, kat_inputs
    for k in HURDLE_KATS:
        d = kat_inputs(k)
        q = k["q_y"] + k["q_zi"]
        betas, gammas, phis, D = map(np.array, (k["betas"], k["gammas"], k["phis"], k["D"]))
        eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], k["family"], nAGQ=k["nAGQ"], X_zi=d["X_zi"], Z_zi=d["Z_zi"])
        gh = gb.GHfun(eng, np.zeros((eng.n, q)), betas, np.linalg.inv(D), phis, gammas)
        assert gh.stats["n_not_converged"] == 0
        tv = np.concatenate([betas, np.log(np.diag(D)) if k["diag_D"] else gb.chol_transf(D), phis, gammas])
        ll = -gb.logLik_mixed(tv, eng, gh, diag_D=k["diag_D"])
        assert abs(ll - k["loglik"]) < k["tol"], (k["name"], ll)
        eng.close()


@pytest.mark.parametrize("fam,link,q_y,q_zi,phis", [
    ("censored.normal", None, 2, 0, [-0.3]), ("binomial", "probit", 2, 0, None), ("binomial", "cloglog", 1, 0, None),
    ("beta.fam", None, 2, 0, [1.5]), ("Gamma.fam", None, 2, 0, [1.0]), ("students.t", None, 1, 0, [-0.4]),
    ("beta.binomial", "cloglog", 2, 0, [1.1]), ("zi.negative.binomial", None, 2, 1, [0.5]),
    ("hurdle.lognormal", None, 1, 1, [-0.3]), ("hurdle.poisson", None, 2, 0, None)], ids=lambda v: str(v))
def test_families_on_clusters_beyond_the_register_variant(fam, link, q_y, q_zi, phis):
    """Clusters of 17 .. 45 observations of the families without a two-pass form: the generic kernel (staged tile
    and tile-outer form) with the careful functors -- which take their table-based branches wherever no link clamp
    is active -- against the oracle."""
    rng = np.random.default_rng(900 + sum(map(ord, fam)) + q_y)
    fam_name = fam
    zi = fam_name.startswith("zi.") or fam_name.startswith("hurdle.")
    sizes = rng.integers(17, 46, 16)
    ids = np.repeat(np.arange(len(sizes)), sizes)
    N = len(ids)
    x = rng.random(N)
    g = rng.integers(0, 2, len(sizes)).astype(float)[ids]
    d = dict(id=ids, X=np.column_stack([np.ones(N), x, g]), Z=np.column_stack([np.ones(N), x])[:, :q_y])
    if zi:
        d["X_zi"] = np.column_stack([np.ones(N), g])
        if q_zi:
            d["Z_zi"] = np.ones((N, 1))
    q = q_y + q_zi
    A = rng.standard_normal((q, q)) * 0.3
    D = A @ A.T + 0.3 * np.eye(q)
    betas = np.array([-0.2, 0.6, -0.3]) if fam in ("beta.fam", "binomial", "beta.binomial") else np.array([0.3, 0.5, -0.4])
    b = rng.standard_normal((len(sizes), q)) @ np.linalg.cholesky(D).T
    eta = d["X"] @ betas + np.sum(d["Z"] * b[ids, :q_y], axis=1)
    d["y"] = response(fam, rng, N, eta)
    th = dict(betas=betas, D=D, phis=None if phis is None else np.array(phis),
              gammas=np.array([-0.8, 0.4]) if zi else None, nAGQ=7 if q >= 3 else 9, df=4.0 if fam == "students.t" else None)
    ofam = oracle_family(fam_name, link)
    od, gh = oracle_grid(d, th, ofam)
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, link=link, nAGQ=th["nAGQ"],
                    X_zi=d.get("X_zi"), Z_zi=d.get("Z_zi"), df=th.get("df"))
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    tv = thetas_vec(th)
    ref = agq.suff_stats(tv, od, ofam, gh)
    got = eng.eval(th["betas"], th["D"], th["phis"], th["gammas"])
    assert got["n_nonfinite"] == ref["n_nonfinite"] == 0
    assert relerr(got["loglik"], ref["loglik"]) < RTOL
    for key in ("g_beta", "g_phi", "g_gamma"):
        if key in ref:
            sc = max(1.0, np.max(np.abs(ref[key])))
            assert np.max(np.abs(got[key] - ref[key])) < 10 * RTOL * sc, key
    assert relerr(got["S"], ref["S"]) < RTOL
    st = eng.find_modes(th["betas"], np.linalg.inv(D), th["phis"], th["gammas"], warm=False)
    assert st["n_chol_failed"] == 0
    modes, chol, logdet = eng.get_modes()
    assert np.max(np.abs(modes - gh.post_modes)) < 1e-7
    eng.close()
