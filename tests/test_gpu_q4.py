"""Random effects of dimension q = 4 (the reference accepts any q, R/mixed_model.R:150-160; the engine compiles
q <= 4): nAGQ^4 nodes run on the generic kernel (tile-outer / streaming forms).  Parity with the oracle to the
north-star tolerance on log-likelihood, score and S; the device mode search against the oracle's Newton
restatement; the one-pass observed information against cd_vec of the score; an E-step / frozen-weight M-step
pair; a whole fit against the oracle's restatement of mixed_fit."""
import numpy as np
import pytest

from oracle import agq
import glmmadaptive_b200 as gb
from helpers import oracle_data, oracle_family, oracle_grid, pack_chol, thetas_vec, relerr
from family_cases import response

pytestmark = pytest.mark.gpu
RTOL = 1e-10

Q4_CASES = [
    # family, link, q_y, q_zi, phis, nAGQ
    ("poisson", None, 4, 0, None, 5),
    ("negative.binomial", None, 4, 0, [0.7], 5),
    ("binomial", "logit", 4, 0, None, 4),
    ("censored.normal", None, 4, 0, [-0.3], 5),
    ("zi.poisson", None, 3, 1, None, 5),
    ("hurdle.negative.binomial", None, 2, 2, [0.6], 5),
    ("poisson", None, 4, 0, None, 7),          # K = 2401: beyond the per-warp node sums, streaming form
]


def q4_case(fam, link, q_y, q_zi, phis, nAGQ, n=24, seed=0):
    rng = np.random.default_rng(1000 + sum(map(ord, fam)) + 10 * q_y + q_zi + seed)
    sizes = 1 + rng.poisson(6, size=n)
    ids = np.repeat(np.arange(n), sizes)
    N = len(ids)
    x = rng.random(N)
    t = rng.random(N)
    g = rng.integers(0, 2, n).astype(float)[ids]
    zi = fam.startswith("zi.") or fam.startswith("hurdle.")
    d = dict(id=ids, X=np.column_stack([np.ones(N), x, g]),
             Z=np.column_stack([np.ones(N), x, t, x * t - 0.25])[:, :q_y])
    if zi:
        d["X_zi"] = np.column_stack([np.ones(N), g])
        if q_zi:
            d["Z_zi"] = np.column_stack([np.ones(N), x])[:, :q_zi]
    q = q_y + q_zi
    A = rng.standard_normal((q, q)) * 0.25
    D = A @ A.T + 0.25 * np.eye(q)
    betas = np.array([-0.2, 0.6, -0.3]) if fam == "binomial" else np.array([0.3, 0.5, -0.4])
    b = rng.standard_normal((n, q)) @ np.linalg.cholesky(D).T
    eta = d["X"] @ betas + np.sum(d["Z"] * b[ids, :q_y], axis=1)
    d["y"] = response(fam, rng, N, eta)
    th = dict(betas=betas, D=D, phis=None if phis is None else np.array(phis),
              gammas=np.array([-0.8, 0.4]) if zi else None, nAGQ=nAGQ)
    return d, th


def _engine(d, th, fam, link):
    return gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam, link=link, nAGQ=th["nAGQ"],
                     X_zi=d.get("X_zi"), Z_zi=d.get("Z_zi"))


@pytest.mark.parametrize("fam,link,q_y,q_zi,phis,nAGQ", Q4_CASES)
def test_q4_parity(fam, link, q_y, q_zi, phis, nAGQ):
    d, th = q4_case(fam, link, q_y, q_zi, phis, nAGQ, n=12 if nAGQ == 7 else 24)
    ofam = oracle_family(fam, link)
    od, gh = oracle_grid(d, th, ofam)
    eng = _engine(d, th, fam, link)
    assert eng.q == 4 and eng.K == nAGQ ** 4
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
    st = eng.find_modes(th["betas"], np.linalg.inv(th["D"]), th["phis"], th["gammas"], warm=False)
    assert st["n_chol_failed"] == 0 and st["n_not_converged"] == 0
    modes, chol, logdet = eng.get_modes()
    assert np.max(np.abs(modes - gh.post_modes)) < 1e-7
    assert np.max(np.abs(logdet - gh.log_dets)) < 1e-7
    eng.close()


def test_q4_observed_information_and_em():
    """q = 4 with a diagonal D (4 D parameters) and with the full D (10): one-pass observed information against
    cd_vec of the score; E-step + frozen-weight scores against the oracle."""
    from glmmadaptive_b200.fit import cd_vec
    d, th = q4_case("poisson", None, 4, 0, None, 5, n=20)
    ofam = oracle_family("poisson")
    od, gh = oracle_grid(d, th, ofam)
    eng = _engine(d, th, "poisson", None)
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    for diag in (True, False):
        thd = dict(th)
        if diag:
            thd["D"] = np.diag(np.diag(th["D"]))
        tv = thetas_vec(thd, diag_D=diag)
        H = eng.observed_information(thd["betas"], thd["D"], None, None, diag_D=diag)
        ref = cd_vec(tv, lambda t: gb.score_mixed(t, eng, None, diag))
        sd = np.sqrt(np.abs(np.diag(ref)))
        assert np.max(np.abs(H - ref) / np.outer(sd, sd)) < 5e-5
    # EM passes (generic kernel, em_mode 1 / 2) against the oracle's restatement of R/mixed_fit.R:119-229
    from oracle import fit as ofit
    betas, D = th["betas"], th["D"]
    p_by, post_b2, ll, eta_zi = ofit.e_step(od, ofam, gh, betas, D, None, None)
    es = eng.em_estep(betas, D, None, None)
    assert relerr(es["loglik"], ll) < RTOL
    assert relerr(es["S"].reshape(-1), np.sum(post_b2, axis=0)) < RTOL
    b2 = betas + np.array([0.05, -0.03, 0.02])
    sc = eng.em_score(b2, None, None)
    ref_b = ofit.score_betas(b2, od, ofam, gh, None, eta_zi, p_by)
    assert np.max(np.abs(-sc["g_beta"] - ref_b)) < 10 * RTOL * max(1.0, np.max(np.abs(ref_b)))
    eng.close()


def test_q4_whole_fit_vs_oracle():
    """mixed_model() with four random effects (nAGQ = 3, K = 81) against the oracle's restatement of mixed_fit."""
    from oracle import fit as ofit
    d, th = q4_case("poisson", None, 4, 0, None, 3, n=60, seed=5)
    ofam = oracle_family("poisson")
    od = oracle_data(d)
    con = dict(nAGQ=3, iter_EM=10, iter_qN_outer=3)
    ref = ofit.mixed_fit(od, ofam, control=con, mode_finder="newton", hessian=False)
    fit = gb.mixed_model(d["y"], d["X"], d["Z"], d["id"], "poisson", control=con, hessian=False)
    assert fit["converged"] == ref["converged"]
    assert np.max(np.abs(fit["coefficients"] - ref["coefficients"])) < 1e-6
    assert np.max(np.abs(fit["D"] - ref["D"])) < 1e-6
    assert abs(fit["logLik"] - ref["logLik"]) < 1e-7 * abs(ref["logLik"])
