"""Published fits of the reference's vignettes for the families of BASELINE configs C2-C5 (poisson,
negative binomial, zero-inflated poisson / negative binomial, beta), transcribed from the rendered
docs/articles/*.html of the reference with line numbers.  The data sets need R's rpois / rgamma /
rnbinom / rbeta, restated in oracle/r_rng.py (RScalarRng) -- the published numbers below are what
pins those restatements AND the oracle's family formulas against real R output.  Shared by the CPU
oracle tests (tests/test_oracle_kat.py) and the GPU engine tests (tests/test_gpu_fit.py).

Each entry: the published estimates (all printed digits) and logLik; `eval_tol` bounds
|logLik(published estimates) - published logLik| (limited by the printed digits of the estimates),
`coef_tol` / `sd_tol` bound |fitted - published| for a whole mixed_model() fit."""
import numpy as np


def _D(sd, corr=None):
    sd = np.asarray(sd, float)
    R = np.eye(len(sd)) if corr is None else np.asarray(corr, float)
    return R * np.outer(sd, sd)


FAMILY_KATS = [
    # gm1 <- mixed_model(y ~ sex * time, random = ~ 1 | id, family = poisson())
    # docs/articles/GLMMadaptive.html:438-492 (summary + cbind(fixef) table)
    dict(name="poisson_gm1", data="poisson", family="poisson", q_y=1, q_zi=0, nAGQ=11, diag_D=False, zi=False,
         betas=[2.94471506, -0.86111218, 0.24039444, -0.05105926], gammas=None, phis=None,
         sd=[0.9781394], D=_D([0.9781394]), loglik=-2782.306, eval_tol=6e-4,
         # this fit stops on the EM criterion far from a stationary point (score 680 on `time`), so the
         # printed digits depend on the mode finder: the reference's BFGS path gives them, Newton 2e-5
         coef_tol=2e-7, sd_tol=2e-7, coef_tol_newton=3e-5, fit_mode_finder="bfgs"),
    # fm1 <- mixed_model(y ~ sex * time, random = ~ 1 | id, family = zi.poisson(), zi_fixed = ~ sex)
    # docs/articles/ZeroInflated_and_TwoPart_Models.html:196-222
    dict(name="zip_fm1", data="zi_count", family="zi.poisson", q_y=1, q_zi=0, nAGQ=11, diag_D=False, zi=True,
         betas=[1.5275644687, 0.0173647260, -0.0040365383, 0.0004326841], gammas=[-1.210597, 0.498611],
         phis=None, sd=[0.8272898], D=_D([0.8272898]), loglik=-2223.937, eval_tol=6e-4,
         coef_tol=2e-10, gamma_tol=6e-7, sd_tol=6e-8),
    # fm2 <- update(fm1, zi_random = ~ 1 | id)                                   html:231-258
    dict(name="zip_fm2", data="zi_count", family="zi.poisson", q_y=1, q_zi=1, nAGQ=11, diag_D=False, zi=True,
         betas=[1.5891884685, -0.0085650817, -0.0030156086, 0.0002487272], gammas=[-1.112309, 0.427672],
         phis=None, sd=[0.7394, 0.6917], D=_D([0.7394, 0.6917], [[1, -0.6550], [-0.6550, 1]]),
         corr=-0.6550, loglik=-2214.277, eval_tol=6e-4, coef_tol=5e-10, gamma_tol=6e-7, sd_tol=6e-5),
    # gm1 <- mixed_model(..., family = zi.negative.binomial(), zi_fixed = ~ sex)  html:279-308
    dict(name="zinb_gm1", data="zi_count", family="zi.negative.binomial", q_y=1, q_zi=0, nAGQ=11, diag_D=False,
         zi=True, betas=[1.467900626, 0.029416703, 0.003005305, -0.006440699], gammas=[-1.5554713, 0.5914944],
         phis=[float(np.log(2.227075))], sd=[0.8293515], D=_D([0.8293515]), loglik=-1958.746, eval_tol=6e-4,
         coef_tol=6e-10, gamma_tol=6e-8, sd_tol=6e-8, phis_tol=3e-7),
    # fm2 <- mixed_model(y ~ f1 * f2, random = ~ 1 | g, family = my_negBinom2(), n_phis = 1,
    #                    initial_values = list(betas = poisson()))  -- the package's negative.binomial()
    # formulas with analytic scores; docs/articles/Custom_Models.html:302-332
    dict(name="negbin_fm2", data="custom_negbin", family="negative.binomial", q_y=1, q_zi=0, nAGQ=11,
         diag_D=False, zi=False, betas=[1.6361832, 0.6031425, 1.0137677, 1.5385349, -0.4378977, -0.5835306],
         gammas=None, phis=[-0.6477284], sd=[0.07476987], D=_D([0.07476987]), loglik=-9994.062, eval_tol=6e-4,
         coef_tol=6e-8, sd_tol=2e-8, phis_tol=6e-8),
    # gm <- mixed_model(y ~ sex * time, random = ~ 1 | id, family = beta.fam(), n_phis = 1)
    # docs/articles/Custom_Models.html:492-517 (the vignette's beta.fam() = the package's, R/Fit_Funs.R:887-930)
    dict(name="beta_gm", data="beta", family="beta.fam", q_y=1, q_zi=0, nAGQ=11, diag_D=False, zi=False,
         betas=[-2.04379396, -0.29655232, 0.21844820, -0.05949062], gammas=None, phis=[1.69036],
         sd=[0.9143138], D=_D([0.9143138]), loglik=581.2972, eval_tol=6e-5,
         coef_tol=6e-9, sd_tol=6e-8, phis_tol=6e-6),
]

_CACHE = {}


def kat_data(k):
    """The regenerated vignette data set of entry k as Engine / agq.Data keyword arguments."""
    from oracle import vignette_data as V
    name = k["data"]
    if name not in _CACHE:
        _CACHE[name] = getattr(V, name + "_data")()
    v = _CACHE[name]
    d = dict(y=v["y"], X=v["X"], Z=v["Z"], id=v["id"])
    if k["zi"]:
        d["X_zi"] = v["X_zi"]
        d["Z_zi"] = v["Z_zi"] if k["q_zi"] else None
    return d


def kat_theta(k):
    arr = lambda v: None if v is None else np.array(v, float)
    return arr(k["betas"]), np.array(k["D"], float), arr(k["phis"]), arr(k["gammas"])


def check_family_fit(k, fit, newton=False):
    ct = k["coef_tol_newton"] if (newton and "coef_tol_newton" in k) else k["coef_tol"]
    st = max(k["sd_tol"], ct) if newton else k["sd_tol"]
    assert fit["converged"], k["name"]
    assert np.max(np.abs(np.asarray(fit["coefficients"]) - np.array(k["betas"]))) < ct, (k["name"], fit["coefficients"])
    if k["gammas"] is not None:
        assert np.max(np.abs(np.asarray(fit["gammas"]) - np.array(k["gammas"]))) < k["gamma_tol"], (k["name"], fit["gammas"])
    if k["phis"] is not None:
        assert np.max(np.abs(np.asarray(fit["phis"]) - np.array(k["phis"]))) < k["phis_tol"], (k["name"], fit["phis"])
    D = np.asarray(fit["D"])
    assert np.max(np.abs(np.sqrt(np.diag(D)) - np.array(k["sd"]))) < st, (k["name"], D)
    if "corr" in k:
        assert abs(D[0, 1] / np.sqrt(D[0, 0] * D[1, 1]) - k["corr"]) < 6e-5, (k["name"], D)
    assert abs(fit["logLik"] - k["loglik"]) < 6e-4, (k["name"], fit["logLik"])
