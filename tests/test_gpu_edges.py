"""GPU: edge cases of the hot path through the C ABI, each against the oracle (1e-10 relative):
single-observation clusters, one very large cluster among tiny ones (ragged CSR, streaming path),
a single cluster, the largest grids the engine accepts (nAGQ = 32 for q = 1; nAGQ = 21 for q = 2,
the vignette's setting), zero cluster weights, offsets in both linear predictors, extreme linear
predictors that activate R's link clamps, and argument errors."""
import ctypes

import numpy as np
import pytest

from oracle import agq
import glmmadaptive_b200 as gb
from glmmadaptive_b200 import _lib
from helpers import oracle_data, oracle_family, oracle_grid, pack_chol, thetas_vec, relerr

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _mk(rng, sizes, q_y, fam, zi=False):
    ids = np.repeat(np.arange(len(sizes)), sizes)
    N = len(ids)
    x = rng.random(N)
    g = rng.integers(0, 2, len(sizes)).astype(float)[ids]
    d = dict(id=ids, X=np.column_stack([np.ones(N), x, g]), Z=np.column_stack([np.ones(N), x])[:, :q_y])
    if zi:
        d["X_zi"] = np.column_stack([np.ones(N), g])
    return d, N


def _check(d, th, fam_name, **kw):
    fam = oracle_family(fam_name)
    od, gh = oracle_grid(d, th, fam)
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, nAGQ=th["nAGQ"], X_zi=d.get("X_zi"),
                    Z_zi=d.get("Z_zi"), offset=d.get("offset"), offset_zi=d.get("offset_zi"),
                    weights=d.get("weights"))
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    ref = agq.suff_stats(thetas_vec(th), od, fam, gh)
    got = eng.eval(th["betas"], th["D"], th.get("phis"), th.get("gammas"))
    assert got["n_nonfinite"] == ref["n_nonfinite"]
    assert relerr(got["loglik"], ref["loglik"]) < RTOL
    for key in ("g_beta", "g_phi", "g_gamma"):
        if key in ref:
            assert np.max(np.abs(got[key] - ref[key])) < 10 * RTOL * max(1.0, np.max(np.abs(ref[key]))), key
    assert np.max(np.abs(got["S"] - ref["S"])) < RTOL * max(1.0, np.max(np.abs(ref["S"])))
    st = eng.find_modes(th["betas"], np.linalg.inv(th["D"]), th.get("phis"), th.get("gammas"), warm=False)
    assert st["n_chol_failed"] == 0
    modes, _, _ = eng.get_modes()
    assert np.max(np.abs(modes - gh.post_modes)) < 1e-7
    eng.close()
    return got, ref


def test_single_observation_clusters_and_single_cluster():
    rng = np.random.default_rng(1)
    betas, D = np.array([0.2, 0.4, -0.3]), np.array([[0.5, 0.1], [0.1, 0.3]])
    d, N = _mk(rng, np.ones(50, int), 2, "poisson")
    d["y"] = rng.poisson(np.exp(d["X"] @ betas)).astype(float)
    _check(d, dict(betas=betas, D=D, nAGQ=11), "poisson")
    d, N = _mk(rng, np.array([7]), 2, "poisson")
    d["y"] = rng.poisson(np.exp(d["X"] @ betas)).astype(float)
    _check(d, dict(betas=betas, D=D, nAGQ=11), "poisson")


def test_one_huge_cluster_among_tiny_ones():
    rng = np.random.default_rng(2)
    sizes = np.array([1, 2, 600, 1, 3, 33, 2, 17, 1, 9])
    betas, D = np.array([0.1, 0.3, -0.2]), np.array([[0.4, 0.05], [0.05, 0.2]])
    d, N = _mk(rng, sizes, 2, "negative.binomial")
    mu = np.exp(d["X"] @ betas)
    d["y"] = rng.negative_binomial(2.0, 2.0 / (2.0 + mu)).astype(float)
    _check(d, dict(betas=betas, D=D, phis=np.array([np.log(2.0)]), nAGQ=9), "negative.binomial")


@pytest.mark.parametrize("q_y,nagq", [(1, 32), (2, 21), (1, 1), (2, 16)])
def test_largest_and_smallest_grids(q_y, nagq):
    rng = np.random.default_rng(3)
    betas = np.array([-0.3, 0.6, -0.2])
    D = np.array([[0.8, 0.1], [0.1, 0.3]])[:q_y, :q_y]
    d, N = _mk(rng, 1 + rng.poisson(4, 25), q_y, "binomial")
    d["y"] = (rng.random(N) < 1 / (1 + np.exp(-(d["X"] @ betas)))).astype(float)
    _check(d, dict(betas=betas, D=D, nAGQ=nagq), "binomial")


def test_zero_weights_offsets_and_zero_part():
    rng = np.random.default_rng(4)
    betas, gammas = np.array([0.4, 0.3, -0.2]), np.array([-1.0, 0.5])
    D = np.array([[0.4, 0.1], [0.1, 0.2]])
    d, N = _mk(rng, 1 + rng.poisson(5, 30), 2, "zi.poisson", zi=True)
    y = rng.poisson(np.exp(d["X"] @ betas)).astype(float)
    y[rng.random(N) < 0.3] = 0.0
    d["y"] = y
    w = rng.uniform(0.5, 2.0, 30)
    w[[3, 11]] = 0.0
    d["weights"] = w
    d["offset"] = rng.normal(0, 0.3, N)
    d["offset_zi"] = rng.normal(0, 0.3, N)
    got, ref = _check(d, dict(betas=betas, D=D, gammas=gammas, nAGQ=11), "zi.poisson")
    assert abs(got["sum_w"] - w.sum()) < 1e-12


def test_link_clamps_on_extreme_linear_predictors():
    """|eta| beyond 30 (logit) and eta below log(eps) (log link): R's make.link clamps are active and
    the register variant must hand these clusters to the careful kernel."""
    rng = np.random.default_rng(5)
    D = np.array([[0.5]])
    d, N = _mk(rng, 1 + rng.poisson(4, 30), 1, "binomial")
    d["X"][:, 1] = rng.normal(0, 25, N)              # wild covariate
    betas = np.array([0.5, 1.5, -0.2])
    d["y"] = (rng.random(N) < 1 / (1 + np.exp(-np.clip(d["X"] @ betas, -40, 40)))).astype(float)
    _check(d, dict(betas=betas, D=D, nAGQ=11), "binomial")
    d2, N2 = _mk(rng, 1 + rng.poisson(4, 30), 1, "poisson")
    d2["X"][:, 1] = rng.normal(0, 1, N2)
    betas2 = np.array([-38.0, 0.5, 0.1])              # exp(eta) < eps: mu clamps to eps
    d2["y"] = np.zeros(N2)
    d2["y"][rng.random(N2) < 0.05] = 1.0
    _check(d2, dict(betas=betas2, D=D, nAGQ=11), "poisson")


def test_argument_errors_are_reported_not_crashed():
    lib = _lib.load()
    rng = np.random.default_rng(6)
    d, N = _mk(rng, np.array([3, 4]), 1, "poisson")
    y = rng.poisson(1.0, N).astype(float)
    with pytest.raises(ValueError):
        gb.Engine(y, d["X"], d["Z"], d["id"], "compoisson")
    with pytest.raises(_lib.GlmmaError):
        gb.Engine(y, d["X"], d["Z"], d["id"], "poisson", nAGQ=99)
    with pytest.raises(_lib.GlmmaError):               # zero part family without X_zi
        gb.Engine(y, d["X"], d["Z"], d["id"], "zi.poisson")
    with pytest.raises(_lib.GlmmaError):               # students.t needs its df
        gb.Engine(y, d["X"], d["Z"], d["id"], "students.t")
    eng = gb.Engine(y, d["X"], d["Z"], d["id"], "poisson")
    with pytest.raises(_lib.GlmmaError) as e:          # eval before any grid exists
        eng.eval(np.zeros(3), np.eye(1))
    assert e.value.code == _lib.ERR_STATE
    eng.find_modes(np.zeros(3), np.eye(1))
    with pytest.raises(_lib.GlmmaError) as e:          # D not positive definite
        eng.eval(np.zeros(3), np.array([[-1.0]]))
    assert e.value.code == _lib.ERR_NUMERIC
    with pytest.raises(_lib.GlmmaError) as e:          # M-step before an E-step
        eng.em_score(np.zeros(3))
    assert e.value.code == _lib.ERR_STATE
    eng.close()


def test_id_labels_of_any_type_and_ungrouped_rows():
    rng = np.random.default_rng(7)
    d, N = _mk(rng, np.array([3, 2, 4]), 1, "poisson")
    y = rng.poisson(1.0, N).astype(float)
    th = (np.array([0.1, 0.2, -0.1]), np.array([[0.5]]))
    ref = None
    for ids in (d["id"], np.array(list("aaabbcccc")), np.repeat([10**12, 5, -3], [3, 2, 4]), np.repeat([2.5, 0.1, 7.0], [3, 2, 4])):
        eng = gb.Engine(y, d["X"], d["Z"], ids, "poisson", nAGQ=7)
        assert eng.n == 3
        eng.find_modes(th[0], np.linalg.inv(th[1]), warm=False)
        r = eng.eval_raw(th[0], th[1])
        ref = r if ref is None else ref
        assert np.array_equal(r, ref)
        eng.close()
    with pytest.raises(ValueError):
        gb.Engine(y, d["X"], d["Z"], np.array([0, 1, 0, 1, 2, 2, 2, 2, 2]), "poisson")


@pytest.mark.parametrize("fam_name", ["negative.binomial", "zi.poisson", "censored.normal"])
def test_all_size_classes_in_one_data_set(fam_name):
    """Cluster sizes 1..40 in one data set: the register variant's three tiles (8, 12, 16 observations,
    each on its own cluster list) and the generic kernel all contribute to one evaluation."""
    rng = np.random.default_rng(11)
    sizes = np.concatenate([rng.integers(1, 9, 40), rng.integers(9, 13, 30), rng.integers(13, 17, 20), rng.integers(17, 41, 10)])
    rng.shuffle(sizes)
    betas, D = np.array([0.2, 0.4, -0.3]), np.array([[0.4, 0.08], [0.08, 0.25]])
    zi = fam_name == "zi.poisson"
    d, N = _mk(rng, sizes, 2, fam_name, zi=zi)
    eta = d["X"] @ betas
    th = dict(betas=betas, D=D, nAGQ=9)
    if fam_name == "negative.binomial":
        d["y"] = rng.negative_binomial(2.0, 2.0 / (2.0 + np.exp(eta))).astype(float)
        th["phis"] = np.array([np.log(2.0)])
    elif zi:
        y = rng.poisson(np.exp(eta)).astype(float)
        y[rng.random(N) < 0.3] = 0.0
        d["y"] = y
        th["gammas"] = np.array([-1.0, 0.5])
    else:
        v = eta + 0.7 * rng.standard_normal(N)
        hi, lo = np.quantile(v, 0.8), np.quantile(v, 0.1)
        d["y"] = np.column_stack([np.clip(v, lo, hi), np.where(v > hi, 2.0, np.where(v < lo, 1.0, 0.0))])
        th["phis"] = np.array([np.log(0.7)])
    d["weights"] = rng.uniform(0.5, 1.5, len(sizes))
    _check(d, th, fam_name)


def test_two_engines_with_different_tiles_on_one_device():
    """Two handles alive on the same device whose launches need different amounts of dynamic shared
    memory (the limit is a per-device property of the kernel function, not of the handle)."""
    rng = np.random.default_rng(12)
    betas, D = np.array([0.2, 0.4, -0.3]), np.array([[0.4, 0.08], [0.08, 0.25]])
    engines, refs = [], []
    for sizes in (np.full(20, 30), np.full(40, 3), np.full(20, 22)):
        d, N = _mk(rng, sizes, 2, "poisson")
        d["y"] = rng.poisson(np.exp(d["X"] @ betas)).astype(float)
        e = gb.Engine(d["y"], d["X"], d["Z"], d["id"], "poisson", nAGQ=9)
        e.find_modes(betas, np.linalg.inv(D), warm=False)
        engines.append(e)
        refs.append(e.eval_raw(betas, D))
    for e, r in zip(engines, refs):          # the earlier handles still launch after the later ones were created
        assert np.array_equal(e.eval_raw(betas, D), r)
        e.close()


@pytest.mark.parametrize("fam_name,q_zi", [("censored.normal", 0), ("zi.poisson", 0), ("zi.poisson", 1),
                                           ("zi.negative.binomial", 1), ("zi.binomial", 0), ("hurdle.poisson", 1),
                                           ("hurdle.negative.binomial", 0), ("hurdle.lognormal", 1),
                                           ("hurdle.beta.fam", 0)])
def test_hoisted_paths_at_their_class_extremes(fam_name, q_zi):
    """The register variant sorts a cluster's observations into classes (censored / uncensored; positive / zero
    response) and hoists the class-wide terms to the node: clusters that consist of ONE class only (all censored,
    none censored, all zeros, no zeros), single observations, and all three tiles, with weights and offsets in
    both linear predictors."""
    rng = np.random.default_rng(31 + q_zi + len(fam_name))
    sizes = np.concatenate([rng.integers(1, 9, 36), rng.integers(9, 13, 12), rng.integers(13, 17, 12)])
    rng.shuffle(sizes)
    n = len(sizes)
    betas, D2 = np.array([0.2, 0.4, -0.3]), np.array([[0.4, 0.08], [0.08, 0.25]])
    zi = fam_name != "censored.normal"
    d, N = _mk(rng, sizes, 2, fam_name, zi=zi)
    ids = d["id"]
    q = 2 + q_zi
    D = D2 if q == 2 else np.array([[0.4, 0.08, -0.05], [0.08, 0.25, 0.03], [-0.05, 0.03, 0.3]])
    if q_zi:
        d["Z_zi"] = np.ones((N, 1))
    d["offset"] = rng.normal(0, 0.2, N)
    if zi:
        d["offset_zi"] = rng.normal(0, 0.3, N)
    d["weights"] = rng.uniform(0.5, 1.5, n)
    eta = d["X"] @ betas + d["offset"]
    th = dict(betas=betas, D=D, nAGQ=7 if q == 3 else 9)
    cls = rng.integers(0, 3, n)[ids]            # per cluster: 0 mixed, 1 all of the first class, 2 all of the second
    if fam_name == "censored.normal":
        v = eta + 0.7 * rng.standard_normal(N)
        ind = rng.integers(0, 3, N).astype(float)
        ind[cls == 1] = 0.0                                        # no observation censored
        ind[cls == 2] = rng.integers(1, 3, int((cls == 2).sum()))  # every observation censored
        d["y"] = np.column_stack([v, ind])
        th["phis"] = np.array([np.log(0.7)])
    else:
        zero = rng.random(N) < 0.35
        zero[cls == 1] = False
        zero[cls == 2] = True
        mu = np.exp(np.clip(eta, -3, 3))
        if fam_name in ("zi.poisson", "zi.negative.binomial"):
            y = rng.poisson(mu).astype(float)
            y[cls == 1] += 1.0
        elif fam_name in ("hurdle.poisson", "hurdle.negative.binomial"):
            y = (1 + rng.poisson(mu)).astype(float)
        elif fam_name == "zi.binomial":
            trials = rng.integers(1, 6, N)
            s = np.maximum(rng.binomial(trials, 1 / (1 + np.exp(-eta))), np.where(cls == 1, 1, 0))
            s[zero] = 0
            y = np.column_stack([s, trials - s]).astype(float)
        elif fam_name == "hurdle.lognormal":
            y = np.exp(eta + 0.5 * rng.standard_normal(N))
        else:
            m = 1 / (1 + np.exp(-eta))
            y = np.clip(rng.beta(m * 5, (1 - m) * 5), 1e-4, 1 - 1e-4)
        if y.ndim == 1:
            y[zero] = 0.0
        d["y"] = y
        th["gammas"] = np.array([-0.8, 0.5])
        if fam_name in ("zi.negative.binomial", "hurdle.negative.binomial"):
            th["phis"] = np.array([0.5])
        elif fam_name == "hurdle.lognormal":
            th["phis"] = np.array([-0.3])
        elif fam_name == "hurdle.beta.fam":
            th["phis"] = np.array([1.5])
    _check(d, th, fam_name)


@pytest.mark.parametrize("fam_name,q_y,big", [("negative.binomial", 2, False), ("poisson", 3, False), ("binomial", 1, False),
                                              ("poisson", 2, False), ("negative.binomial", 2, True), ("binomial", 2, True),
                                              ("poisson", 1, True)])
def test_two_pass_tile_of_the_register_variant(fam_name, q_y, big):
    """Clusters of 17 .. 32 observations only (big: 33 .. 150, i.e. several records -- tiles of 32 rows for pass A,
    then for pass B -- in one warp's queue): the register variant's two-pass form (tile 32, carrier families) is the
    only kernel that runs.  Fused evaluation, E-step / frozen-weight M-step and the per-cluster / per-observation
    contributions against the oracle, with weights and an offset."""
    from oracle import fit as ofit
    rng = np.random.default_rng(50 + q_y + len(fam_name) + (7 if big else 0))
    sizes = rng.integers(33, 151, 24) if big else rng.integers(17, 33, 40)
    sizes[:3] = (33, 64, 97) if big else (17, 32, 25)
    betas = np.array([0.2, 0.4, -0.3])
    D = {1: np.array([[0.5]]), 2: np.array([[0.4, 0.08], [0.08, 0.25]]),
         3: np.array([[0.4, 0.08, -0.05], [0.08, 0.25, 0.03], [-0.05, 0.03, 0.2]])}[q_y]
    ids = np.repeat(np.arange(len(sizes)), sizes)
    N = len(ids)
    x = rng.random(N)
    g = rng.integers(0, 2, len(sizes)).astype(float)[ids]
    d = dict(id=ids, X=np.column_stack([np.ones(N), x, g]), Z=np.column_stack([np.ones(N), x, x * x - 0.3])[:, :q_y])
    d["offset"] = rng.normal(0, 0.2, N)
    d["weights"] = rng.uniform(0.5, 1.5, len(sizes))
    eta = d["X"] @ betas + d["offset"]
    th = dict(betas=betas, D=D, nAGQ=7 if q_y == 3 else 11)
    if fam_name == "negative.binomial":
        d["y"] = rng.negative_binomial(2.0, 2.0 / (2.0 + np.exp(eta))).astype(float)
        th["phis"] = np.array([np.log(2.0)])
    elif fam_name == "poisson":
        d["y"] = rng.poisson(np.exp(eta)).astype(float)
    else:
        d["y"] = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(float)
    _check(d, th, fam_name)
    fam = oracle_family(fam_name)
    od, gh = oracle_grid(d, th, fam)
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, nAGQ=th["nAGQ"], offset=d["offset"], weights=d["weights"])
    eng.set_modes(gh.post_modes, pack_chol(gh.chol_hessians))
    phis = th.get("phis")
    p_by, post_b2, ll, eta_zi = ofit.e_step(od, fam, gh, betas, D, phis, None)
    es = eng.em_estep(betas, D, phis, None)
    assert relerr(es["loglik"], ll) < RTOL
    b2 = betas + np.array([0.05, -0.03, 0.02])
    ph2 = None if phis is None else phis + 0.07
    sc = eng.em_score(b2, ph2, None)
    ref_b = ofit.score_betas(b2, od, fam, gh, ph2, eta_zi, p_by)
    assert np.max(np.abs(-sc["g_beta"] - ref_b)) < 10 * RTOL * max(1.0, np.max(np.abs(ref_b)))
    # per-cluster / per-observation contributions (i_contributions = TRUE)
    tvf = thetas_vec(th)
    ll_i = gb.logLik_mixed(tvf, eng, None, i_contributions=True)
    assert relerr(ll_i, agq.logLik_mixed(tvf, od, fam, gh, i_contributions=True)) < RTOL
    cr = agq.score_mixed(tvf, od, fam, gh, i_contributions=True)
    cg = gb.score_mixed(tvf, eng, None, i_contributions=True)
    w_obs = d["weights"][d["id"]]
    assert np.max(np.abs(cg["score_betas"] - cr["score_betas"] * w_obs[:, None])) < 1e-9
    eng.close()


@pytest.mark.parametrize("fam_name,p,p_zi", [("negative.binomial", 14, 0), ("zi.poisson", 12, 7), ("censored.normal", 20, 0),
                                             ("poisson", 27, 0)])
def test_many_fixed_effect_columns(fam_name, p, p_zi):
    """Wide design matrices: the register variant stages p + p_zi design columns per cluster and folds X'zbar into
    the block partials (up to 32 columns); sizes 1 .. 40 so that every tile and the large-cluster paths see them."""
    rng = np.random.default_rng(70 + p + p_zi)
    sizes = np.concatenate([rng.integers(1, 9, 30), rng.integers(9, 17, 20), rng.integers(17, 41, 8)])
    rng.shuffle(sizes)
    ids = np.repeat(np.arange(len(sizes)), sizes)
    N = len(ids)
    X = np.column_stack([np.ones(N), rng.normal(0, 0.5, (N, p - 1))])
    Z = np.column_stack([np.ones(N), X[:, 1]])
    betas = np.concatenate([[0.2], rng.normal(0, 0.15, p - 1)])
    D = np.array([[0.4, 0.08], [0.08, 0.25]])
    d = dict(id=ids, X=X, Z=Z)
    eta = X @ betas
    th = dict(betas=betas, D=D, nAGQ=9)
    if fam_name == "negative.binomial":
        d["y"] = rng.negative_binomial(2.0, 2.0 / (2.0 + np.exp(eta))).astype(float)
        th["phis"] = np.array([np.log(2.0)])
    elif fam_name == "poisson":
        d["y"] = rng.poisson(np.exp(eta)).astype(float)
    elif fam_name == "zi.poisson":
        d["X_zi"] = np.column_stack([np.ones(N), rng.normal(0, 0.5, (N, p_zi - 1))])
        y = rng.poisson(np.exp(eta)).astype(float)
        y[rng.random(N) < 0.3] = 0.0
        d["y"] = y
        th["gammas"] = np.concatenate([[-1.0], rng.normal(0, 0.2, p_zi - 1)])
    else:
        v = eta + 0.7 * rng.standard_normal(N)
        hi, lo = np.quantile(v, 0.8), np.quantile(v, 0.1)
        d["y"] = np.column_stack([np.clip(v, lo, hi), np.where(v > hi, 2.0, np.where(v < lo, 1.0, 0.0))])
        th["phis"] = np.array([np.log(0.7)])
    _check(d, th, fam_name)


@pytest.mark.parametrize("fam_name,nagq", [("binomial", 11), ("poisson", 16), ("negative.binomial", 7), ("binomial", 3)])
def test_pair_tile_random_intercepts(fam_name, nagq, monkeypatch):
    """q = 1: two adjacent clusters of at most 8 observations share a warp (lanes 0..15 / 16..31 are their nodes).
    Mixed sizes so that pairs, unpaired small clusters and larger clusters all occur; against the oracle, and
    against the one-cluster-per-warp path (GLMMA_NO_PAIR) for the contributions and the EM passes."""
    rng = np.random.default_rng(80 + nagq)
    sizes = np.concatenate([rng.integers(1, 9, 60), rng.integers(9, 30, 8), rng.integers(1, 9, 31)])
    rng.shuffle(sizes)
    betas, D = np.array([0.2, 0.4, -0.3]), np.array([[0.6]])
    d, N = _mk(rng, sizes, 1, fam_name)
    d["offset"] = rng.normal(0, 0.2, N)
    d["weights"] = rng.uniform(0.5, 1.5, len(sizes))
    eta = d["X"] @ betas + d["offset"]
    th = dict(betas=betas, D=D, nAGQ=nagq)
    if fam_name == "negative.binomial":
        d["y"] = rng.negative_binomial(2.0, 2.0 / (2.0 + np.exp(eta))).astype(float)
        th["phis"] = np.array([np.log(2.0)])
    elif fam_name == "poisson":
        d["y"] = rng.poisson(np.exp(eta)).astype(float)
    else:
        d["y"] = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(float)
    _check(d, th, fam_name)
    out = {}
    for tag in ("pair", "single"):
        if tag == "single":
            monkeypatch.setenv("GLMMA_NO_PAIR", "1")
        eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam_name, nAGQ=nagq, offset=d["offset"], weights=d["weights"])
        eng.find_modes(betas, np.linalg.inv(D), th.get("phis"), None, warm=False)
        c = eng.eval_contrib(betas, D, th.get("phis"), None)
        es = eng.em_estep(betas, D, th.get("phis"), None)
        ms = eng.em_score(betas + 0.03, th.get("phis"), None)
        out[tag] = (eng.eval_raw(betas, D, th.get("phis"), None), c, es, ms)
        eng.close()
    monkeypatch.delenv("GLMMA_NO_PAIR")
    (r1, c1, e1, m1), (r2, c2, e2, m2) = out["pair"], out["single"]
    assert np.max(np.abs(r1 - r2)) < 1e-13 * np.max(np.abs(r2))
    for k in ("ll_i", "zbar", "S_i"):
        assert np.max(np.abs(c1[k] - c2[k])) <= 1e-12 * max(1.0, np.max(np.abs(c2[k]))), k
    assert abs(e1["loglik"] - e2["loglik"]) < 1e-13 * abs(e2["loglik"])
    assert np.max(np.abs(m1["g_beta"] - m2["g_beta"])) < 1e-11 * max(1.0, np.max(np.abs(m2["g_beta"])))
