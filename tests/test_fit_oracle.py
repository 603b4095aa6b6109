"""CPU: pins the oracle's restatement of the reference's fitting loop (oracle/fit.py: glm.fit starts,
EM phase, nearPD, optim's BFGS, R/mixed_fit.R) on the coefficients the reference's rendered vignette
publishes for `fm1 <- mixed_model(y ~ sex * time, random = ~ 1 | id, family = binomial())` at
nAGQ = 11 / 15 / 21 (docs/articles/GLMMadaptive.html:332-338 of the reference),
and checks the host helpers of the product (glmmadaptive_b200/fit.py) against the
oracle's independent versions."""
import json
import os

import numpy as np
import pytest

from oracle import agq, families as F, vignette_data, fit as ofit
from glmmadaptive_b200 import fit as mm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vignette_kat.json")


@pytest.mark.parametrize("name", ["D1_nAGQ11", "D1_nAGQ15", "D1_nAGQ21"])
def test_oracle_fit_reproduces_published_fm1(name):
    k = [r for r in json.load(open(GOLD))["kats"] if r["name"] == name][0]
    v = vignette_data.binary_data()
    data = agq.Data(y=v["y"], X=v["X"], Z=v["Z"][:, :1], id=v["id"])
    out = ofit.mixed_fit(data, F.binomial(), control=dict(nAGQ=k["nAGQ"]), hessian=False)
    assert out["converged"]
    # every digit the vignette prints (8 decimals at nAGQ 11 / 15, 7 at nAGQ 21)
    tol = 6e-8 if k["nAGQ"] == 21 else 1.5e-8
    assert np.max(np.abs(out["coefficients"] - np.array(k["betas"]))) < tol
    assert abs(out["D"][0, 0] - k["D"][0][0]) < tol
    assert abs(out["logLik"] - k["loglik"]) < 2e-7


def test_glm_fit_product_vs_oracle():
    rng = np.random.default_rng(0)
    X = np.column_stack([np.ones(300), rng.normal(size=300), rng.integers(0, 2, 300)])
    eta = X @ np.array([0.2, 0.5, -0.4])
    yb = (rng.random(300) < 1 / (1 + np.exp(-eta))).astype(float)
    yp = rng.poisson(np.exp(eta)).astype(float)
    yg = rng.gamma(3.0, 1.0 / (3.0 * (1.5 + 0.2 * X[:, 1])))
    y2 = np.column_stack([rng.binomial(5, 1 / (1 + np.exp(-eta))), np.zeros(300)])
    y2[:, 1] = 5 - y2[:, 0]
    off = rng.normal(0, 0.2, 300)
    for fam, y, o in (("binomial", yb, None), ("poisson", yp, off), ("Gamma", yg, None), ("binomial", y2, None),
                      ("gaussian", eta + rng.normal(size=300), None)):
        a, b = mm.glm_fit(X, y, fam, o), ofit.glm_fit(X, y, fam, o)
        assert np.max(np.abs(a - b)) < 1e-9, fam
    # score equations hold at the solution (poisson with offset)
    b = mm.glm_fit(X, yp, "poisson", off)
    assert np.max(np.abs(X.T @ (yp - np.exp(X @ b + off)))) < 1e-6


def test_nearpd_fdvec_vmmin_product_vs_oracle():
    rng = np.random.default_rng(1)
    A = rng.normal(size=(5, 5))
    M = A + A.T                                   # indefinite
    P1, P2 = mm.nearPD(M), ofit.nearPD(M)
    assert np.max(np.abs(P1 - P2)) < 1e-12
    assert np.linalg.eigvalsh(P1).min() > 0
    Mpd = A @ A.T + np.eye(5)
    assert np.max(np.abs(mm.nearPD(Mpd) - Mpd)) < 1e-10
    f = lambda x: np.array([np.sin(x[0]) * x[1], x[0] * x[1] ** 2, x[2] ** 3 + x[0]])
    x0 = np.array([0.3, -1.7, 2.0])
    assert np.array_equal(mm.fd_vec(x0, f), ofit.fd_vec(x0, f))
    assert np.array_equal(mm.cd_vec(x0, f), agq.cd_vec(x0, f))
    # optim's BFGS on Rosenbrock: identical iterates in both restatements
    fn = lambda x: 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2
    gr = lambda x: np.array([-400 * x[0] * (x[1] - x[0] ** 2) - 2 * (1 - x[0]), 200 * (x[1] - x[0] ** 2)])
    for maxit in (10, 100):
        p1, v1, c1, nf, ng = mm.vmmin(np.array([-1.2, 1.0]), fn, gr, maxit, 1e-8, np.array([1.0, 0.5]))
        o = agq.vmmin(np.array([-1.2, 1.0]), fn, gr, maxit=maxit, reltol=1e-8, parscale=np.array([1.0, 0.5]))
        assert np.array_equal(p1, o["par"]) and v1 == o["value"] and c1 == o["convergence"]
        assert (nf, ng) == tuple(o["counts"])
    assert np.max(np.abs(p1 - 1.0)) < 1e-3      # R documents optim(BFGS) reaching (0.9998, 0.9996) on this problem


def test_starting_values_product_vs_oracle():
    v = vignette_data.hurdle_count_data()
    data = agq.Data(y=v["y"], X=v["X"], Z=v["Z"], id=v["id"], X_zi=v["X_zi"], Z_zi=v["Z_zi"])
    fam = F.hurdle_poisson()
    a = mm.starting_values(v["y"], v["X"], v["X_zi"], fam.family, 3, 0)
    b = ofit.initial_values(data, fam)
    for key in ("betas", "gammas", "D"):
        assert np.max(np.abs(np.asarray(a[key]) - np.asarray(b[key]))) < 1e-9


@pytest.mark.parametrize("name", ["km1", "km2", "hm1", "hm2"])
def test_oracle_fit_reproduces_published_two_part_fits(name):
    """Whole fits of the two-part vignette (hurdle log-normal q = 1 and q = 2 with a diagonal D, hurdle
    negative binomial q = 2 and q = 3): glm.fit starts for betas and gammas, phis, the EM phase and
    optim's BFGS all have to be right for the oracle to print the vignette's digits."""
    from kat_hurdle import FIT_KATS, kat_inputs, check_fit_against_kat
    from helpers import oracle_family
    k = [r for r in FIT_KATS if r["name"] == name][0]
    data = agq.Data(**kat_inputs(k))
    out = ofit.mixed_fit(data, oracle_family(k["family"]), control=dict(nAGQ=k["nAGQ"]), diag_D=k["diag_D"], hessian=False)
    check_fit_against_kat(k, out)


# summary(fm) of docs/articles/Multiple_Comparisons.html:60-105 of the reference: estimates, STANDARD
# ERRORS (sqrt(diag(solve(Hessian))), the cd_vec Hessian of score_mixed), random-intercept sd, logLik
MULTCOMP_FM = dict(est=[-2.2865, 0.8860, 0.3066, -0.5053, 0.5632], se=[0.2496, 0.2447, 0.2266, 0.2463, 0.2237],
                   sd=1.447826, loglik=-583.4797)


def check_multcomp(out):
    k = MULTCOMP_FM
    assert np.max(np.abs(np.asarray(out["coefficients"]) - np.array(k["est"]))) < 6e-5
    se = np.sqrt(np.diag(np.linalg.inv(out["Hessian"])))[:5]
    assert np.max(np.abs(se - np.array(k["se"]))) < 6e-5
    assert abs(np.sqrt(out["D"][0, 0]) - k["sd"]) < 6e-7
    assert abs(out["logLik"] - k["loglik"]) < 6e-5


def test_oracle_fit_reproduces_published_standard_errors():
    v = vignette_data.multcomp_binary_data()
    data = agq.Data(y=v["y"], X=v["X"], Z=v["Z"], id=v["id"])
    check_multcomp(ofit.mixed_fit(data, F.binomial(), hessian=True))


def _fam_kats():
    from kat_families import FAMILY_KATS
    return FAMILY_KATS


@pytest.mark.parametrize("name", [k["name"] for k in _fam_kats()])
def test_oracle_fit_reproduces_published_family_fits(name):
    """Whole fits on the regenerated vignette data land on every digit the reference's rendered
    vignettes print for poisson, zi.poisson (q = 1 and 2), zi.negative.binomial, negative binomial and
    beta -- this pins the oracle's log_dens / score_eta / score_phis / score_eta_zi of those families,
    the EM and quasi-Newton phases and the rpois / rgamma / rnbinom / rbeta restatements on R output."""
    import warnings
    from kat_families import kat_data, check_family_fit
    from helpers import oracle_family
    k = [r for r in _fam_kats() if r["name"] == name][0]
    data = agq.Data(**kat_data(k))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = ofit.mixed_fit(data, oracle_family(k["family"]), control=dict(nAGQ=k["nAGQ"]), hessian=False,
                             mode_finder=k.get("fit_mode_finder", "newton"))
    check_family_fit(k, out)


def test_oracle_penalized_fit_reproduces_published_gm2():
    """gm2 <- mixed_model(y ~ sex * time, random = ~ 1 | id, family = poisson(), penalized = TRUE):
    cbind(fixef(gm1), fixef(gm2)) of docs/articles/GLMMadaptive.html:471-478 prints the penalised
    estimates with seven decimals; the oracle's fit (reference-faithful BFGS mode finder, as gm1 in
    tests/kat_families.py) lands on them."""
    import warnings
    v = vignette_data.poisson_data()
    data = agq.Data(y=v["y"], X=v["X"], Z=v["Z"], id=v["id"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = ofit.mixed_fit(data, F.poisson(), hessian=False, mode_finder="bfgs", penalized=True)
    assert out["converged"]
    assert np.max(np.abs(out["coefficients"] - np.array([2.9433925, -0.8599374, 0.2404593, -0.0511505]))) < 5e-7
