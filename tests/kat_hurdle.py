"""Published fits of the reference's two-part vignette (rendered docs/articles/
ZeroInflated_and_TwoPart_Models.html of the reference, transcribed with line numbers), shared by the
oracle KAT tests (CPU) and the engine KAT tests (GPU).  Standard deviations / correlations are printed
with four digits, hence the 5e-3 tolerance on logLik."""
import numpy as np


def _D(sd, corr=None):
    sd = np.asarray(sd, float)
    R = np.eye(len(sd)) if corr is None else np.asarray(corr, float)
    return R * np.outer(sd, sd)


HURDLE_KATS = [
    # km2 <- update(km1, random = ~ 1 || id, zi_random = ~ 1 | id)          html:496-523
    dict(name="km2", data="hurdle_lognormal", family="hurdle.lognormal", q_y=1, q_zi=1, nAGQ=11, diag_D=True,
         betas=[-2.0811898, -0.2405682, 0.2230215, -0.0568855], gammas=[-1.7160430, 0.7096562], phis=[-0.3349759],
         D=_D([0.9609, 0.6043]), loglik=-1210.313, tol=5e-3),
    # hm1 <- mixed_model(y ~ sex * time, random = ~ time | id, family = hurdle.negative.binomial(), zi_fixed = ~ sex)   html:727-757
    dict(name="hm1", data="hurdle_count", family="hurdle.negative.binomial", q_y=2, q_zi=0, nAGQ=11, diag_D=False,
         betas=[1.48250218, 0.07508276, 0.10249953, -0.08978082], gammas=[-1.4178427, 0.4857483], phis=[0.6778918],
         D=_D([0.7305, 0.3249], [[1, -0.1170], [-0.1170, 1]]), loglik=-2181.468, tol=5e-3),
    # hm2 <- update(hm1, zi_random = ~ 1 | id)                               html:767-799
    dict(name="hm2", data="hurdle_count", family="hurdle.negative.binomial", q_y=2, q_zi=1, nAGQ=7, diag_D=False,
         betas=[1.47627667, 0.07040678, 0.10535938, -0.08613120], gammas=[-1.5420047, 0.5197229], phis=[0.6759598],
         D=_D([0.7315, 0.3258, 0.6598], [[1, -0.1155, -0.0517], [-0.1155, 1, 0.1412], [-0.0517, 0.1412, 1]]),
         loglik=-2176.312, tol=5e-3),
]


def kat_inputs(k):
    from oracle import vignette_data
    v = vignette_data.hurdle_lognormal_data() if k["data"] == "hurdle_lognormal" else vignette_data.hurdle_count_data()
    return dict(y=v["y"], X=v["X"], Z=v["Z"][:, :k["q_y"]], id=v["id"], X_zi=v["X_zi"],
                Z_zi=v["Z_zi"] if k["q_zi"] else None)


# Published ESTIMATES of whole mixed_model() fits on the same data (same html file): the fitting loop
# (glm.fit starts, EM phase, optim BFGS) has to land on these digits.  coef_tol: the digits printed
# (dm2's path has three outer iterations and depends on the mode finder at the 5e-6 level).
FIT_KATS = [
    dict(name="km1", data="hurdle_lognormal", family="hurdle.lognormal", q_y=1, q_zi=0, diag_D=False, nAGQ=11,
         betas=[-2.08122468, -0.24046281, 0.22303981, -0.05686286], gammas=[-1.6034499, 0.6713555], phis=[-0.3348787],
         sd=[0.9609426], loglik=-1214.208, coef_tol=2e-7, sd_tol=2e-7),                       # html:458-489
    dict(name="km2", data="hurdle_lognormal", family="hurdle.lognormal", q_y=1, q_zi=1, diag_D=True, nAGQ=11,
         betas=[-2.0811898, -0.2405682, 0.2230215, -0.0568855], gammas=[-1.7160430, 0.7096562], phis=[-0.3349759],
         sd=[0.9609, 0.6043], loglik=-1210.313, coef_tol=2e-7, sd_tol=6e-5),
    dict(name="hm1", data="hurdle_count", family="hurdle.negative.binomial", q_y=2, q_zi=0, diag_D=False, nAGQ=11,
         betas=[1.48250218, 0.07508276, 0.10249953, -0.08978082], gammas=[-1.4178427, 0.4857483], phis=[0.6778918],
         sd=[0.7305, 0.3249], loglik=-2181.468, coef_tol=2e-7, sd_tol=6e-5),
    dict(name="hm2", data="hurdle_count", family="hurdle.negative.binomial", q_y=2, q_zi=1, diag_D=False, nAGQ=7,
         betas=[1.47627667, 0.07040678, 0.10535938, -0.08613120], gammas=[-1.5420047, 0.5197229], phis=[0.6759598],
         sd=[0.7315, 0.3258, 0.6598], loglik=-2176.312, coef_tol=2e-7, sd_tol=6e-5),
    dict(name="dm2", data="hurdle_count", family="hurdle.poisson", q_y=2, q_zi=1, diag_D=False, nAGQ=7,
         betas=[1.52792269, 0.08389745, 0.09206185, -0.08020395], gammas=[-1.5398382, 0.5175922], phis=None,
         sd=[0.8226, 0.3644, 0.6584], loglik=-2797.307, coef_tol=1e-5, sd_tol=6e-5),               # html:682-706
]


def check_fit_against_kat(k, fit):
    assert fit["converged"], k["name"]
    assert np.max(np.abs(np.asarray(fit["coefficients"]) - np.array(k["betas"]))) < k["coef_tol"], (k["name"], fit["coefficients"])
    assert np.max(np.abs(np.asarray(fit["gammas"]) - np.array(k["gammas"]))) < k["coef_tol"], (k["name"], fit["gammas"])
    if k["phis"] is not None:
        assert np.max(np.abs(np.asarray(fit["phis"]) - np.array(k["phis"]))) < k["coef_tol"], (k["name"], fit["phis"])
    assert np.max(np.abs(np.sqrt(np.diag(fit["D"])) - np.array(k["sd"]))) < k["sd_tol"], (k["name"], fit["D"])
    assert abs(fit["logLik"] - k["loglik"]) < 6e-4, (k["name"], fit["logLik"])
