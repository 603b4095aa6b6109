"""Shared helpers of the parity tests: oracle-side grids and sufficient statistics
for the data/theta dictionaries produced by glmmadaptive_b200.synth / the vignette
data, and packing of the oracle's Cholesky factors in the C ABI's layout."""
import numpy as np

from oracle import agq, families as F

ORACLE_FAMILIES = {
    "binomial": F.binomial, "poisson": F.poisson, "negative.binomial": F.negative_binomial,
    "zi.poisson": F.zi_poisson, "zi.negative.binomial": F.zi_negative_binomial,
    "zi.binomial": F.zi_binomial, "hurdle.poisson": F.hurdle_poisson,
    "hurdle.negative.binomial": F.hurdle_negative_binomial, "hurdle.lognormal": F.hurdle_lognormal,
    "beta.fam": F.beta_fam, "hurdle.beta.fam": F.hurdle_beta_fam, "Gamma.fam": F.Gamma_fam,
    "censored.normal": F.censored_normal, "students.t": lambda: F.students_t(TEST_DF),
    "beta.binomial": F.beta_binomial,
}
TEST_DF = 4.0      # df of the students.t family object used throughout the tests


def oracle_data(d):
    return agq.Data(y=d["y"], X=d["X"], Z=d["Z"], id=d["id"], X_zi=d.get("X_zi"), Z_zi=d.get("Z_zi"),
                    offset=d.get("offset"), offset_zi=d.get("offset_zi"), weights=d.get("weights"))


def oracle_family(name, link=None):
    cls = ORACLE_FAMILIES[name]
    return cls(link) if (name in ("binomial", "beta.binomial") and link) else cls()


def pack_chol(chol_list):
    """list of upper factors -> n x q(q+1)/2 in R's U[upper.tri(U, TRUE)] order."""
    q = chol_list[0].shape[0]
    idx = [(r, c) for c in range(q) for r in range(c + 1)]
    return np.array([[R[r, c] for r, c in idx] for R in chol_list])


def thetas_vec(theta, diag_D=False):
    parts = [theta["betas"]]
    D = np.asarray(theta["D"], float)
    parts.append(np.log(np.diag(D)) if diag_D else agq.chol_transf(D))
    if theta.get("phis") is not None:
        parts.append(theta["phis"])
    if theta.get("gammas") is not None:
        parts.append(theta["gammas"])
    return np.concatenate([np.atleast_1d(np.asarray(p, float)) for p in parts])


def oracle_grid(d, theta, fam, mode_finder="newton", start=None):
    od = oracle_data(d)
    q = od.nRE
    b0 = np.zeros((od.n, q)) if start is None else start
    invD = np.linalg.inv(np.asarray(theta["D"], float))
    gh = agq.GHfun(b0, od, fam, theta["betas"], invD, theta.get("phis"), theta.get("gammas"),
                   theta["nAGQ"], q, mode_finder=mode_finder)
    return od, gh


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) if a.size else 0.0
