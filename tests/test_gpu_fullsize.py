"""GPU: BASELINE.json's full-size workload (config 3: negative binomial, q = 2, nAGQ = 11, 1e6 clusters
x 8 observations) checked through size-independent properties, since the oracle cannot run at this
size (its N x K temporaries are 7.7 GB each):

  * determinism: repeated evaluations are bit-identical (optim's line search compares values);
  * additivity over clusters: the evaluation of the whole set equals the sum over two disjoint
    shards evaluated by separate engines (what the multi-GPU path relies on), to 1e-12;
  * subsample = oracle: the first 300 clusters, evaluated as their own engine on the grid the full
    engine found for them, match the oracle to 1e-10 -- ties the full-size run to the reference
    arithmetic;
  * linearity in the weights: doubling every cluster weight doubles logLik, score and S exactly;
  * the score is the gradient: central differences of the full-size logLik in beta, phi and the
    D parameters match score_mixed to 1e-6 relative (the grid is fixed, as for optim);
  * the engine's modes are stationary points: a warm restart moves no mode by more than 1e-9.
"""
import numpy as np
import pytest

from oracle import agq
import glmmadaptive_b200 as gb
from glmmadaptive_b200 import synth
from helpers import oracle_data, oracle_family, pack_chol, thetas_vec, relerr

pytestmark = pytest.mark.gpu
N_FULL = 1_000_000


@pytest.fixture(scope="module")
def full():
    d, th = synth.make("C3", n=N_FULL, seed=None)
    fam = d.pop("family")
    eng = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam, nAGQ=th["nAGQ"])
    st = eng.find_modes(th["betas"], np.linalg.inv(th["D"]), th["phis"], None, warm=False)
    assert st["n_not_converged"] == 0 and st["n_chol_failed"] == 0
    yield d, th, fam, eng
    eng.close()


def test_determinism_and_counts(full):
    d, th, fam, eng = full
    a = eng.eval_raw(th["betas"], th["D"], th["phis"], None)
    b = eng.eval_raw(th["betas"], th["D"], th["phis"], None)
    assert np.array_equal(a, b)
    r = eng.unpack(a)
    assert r["n_nonfinite"] == 0 and r["sum_w"] == N_FULL and np.isfinite(r["loglik"])


def test_additivity_over_cluster_shards(full):
    d, th, fam, eng = full
    whole = eng.eval_raw(th["betas"], th["D"], th["phis"], None)
    modes, chol, _ = eng.get_modes()
    cut_c = 400_003                       # an uneven split
    cut_r = cut_c * 8
    parts = []
    for cs, rs in ((slice(0, cut_c), slice(0, cut_r)), (slice(cut_c, None), slice(cut_r, None))):
        e = gb.Engine(d["y"][rs], d["X"][rs], d["Z"][rs], d["id"][rs], fam, nAGQ=th["nAGQ"])
        e.set_modes(modes[cs], chol[cs])
        parts.append(e.eval_raw(th["betas"], th["D"], th["phis"], None))
        e.close()
    tot = parts[0] + parts[1]
    # logLik, sum_w and S are sums of same-signed terms: relative 1e-12.  The score components are sums
    # of 8e6 terms of order one that nearly cancel at the data-generating theta: their agreement is
    # limited by the summation order, i.e. by 1e-16 x (sum of |terms|), not by their own magnitude
    rel = np.abs(tot - whole) / np.maximum(np.abs(whole), 1.0)
    assert rel[0] < 1e-12 and rel[1] == 0 and np.max(rel[8:]) < 1e-12
    assert np.max(np.abs(tot - whole)[4:8]) < 1e-15 * 8 * N_FULL * 100


def test_subsample_matches_oracle(full):
    d, th, fam, eng = full
    k = 300
    r1 = k * 8
    sub = dict(y=d["y"][:r1], X=d["X"][:r1], Z=d["Z"][:r1], id=d["id"][:r1])
    modes, chol, logdet = eng.get_modes()
    od = oracle_data(sub)
    ofam = oracle_family(fam)
    q = 2
    idx = [(r, c) for c in range(q) for r in range(c + 1)]
    hess = []
    for row in chol[:k]:
        R = np.zeros((q, q))
        for v, (r, c) in zip(row, idx):
            R[r, c] = v
        hess.append(R.T @ R)
    gh = agq.GHfun_from_modes(od, modes[:k], hess, th["nAGQ"], q)
    ref = agq.suff_stats(thetas_vec(th), od, ofam, gh)
    e = gb.Engine(sub["y"], sub["X"], sub["Z"], sub["id"], fam, nAGQ=th["nAGQ"])
    e.set_modes(modes[:k], chol[:k])
    got = e.eval(th["betas"], th["D"], th["phis"], None)
    e.close()
    assert relerr(got["loglik"], ref["loglik"]) < 1e-10
    assert np.max(np.abs(got["g_beta"] - ref["g_beta"])) < 1e-9 * max(1.0, np.max(np.abs(ref["g_beta"])))
    assert relerr(got["S"], ref["S"]) < 1e-10
    # and the engine's own modes for these clusters are the oracle's Newton modes
    om, _, _ = agq.find_modes_newton(np.zeros((k, q)), od, ofam, th["betas"], np.linalg.inv(th["D"]), th["phis"], None)
    assert np.max(np.abs(om - modes[:k])) < 1e-8


def test_weight_linearity(full):
    d, th, fam, eng = full
    base = eng.eval_raw(th["betas"], th["D"], th["phis"], None)
    modes, chol, _ = eng.get_modes()
    e = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam, nAGQ=th["nAGQ"], weights=np.full(N_FULL, 2.0))
    e.set_modes(modes, chol)
    dbl = e.eval_raw(th["betas"], th["D"], th["phis"], None)
    e.close()
    ref = 2.0 * base
    ref[2] = base[2]                      # the count of dropped clusters is not weighted
    assert np.array_equal(dbl, ref)


def test_score_is_gradient_at_full_size(full):
    d, th, fam, eng = full
    tv = thetas_vec(th)
    # move away from the data-generating theta so that the score is not ~0 relative to its scale
    tv = tv + np.array([0.02, -0.03, 0.01, 0.02, -0.01, 0.015, 0.03])
    sc = gb.score_mixed(tv, eng, None)
    for i in range(len(tv)):
        h = 1e-5
        up, dn = tv.copy(), tv.copy()
        up[i] += h
        dn[i] -= h
        fd = (gb.logLik_mixed(up, eng, None) - gb.logLik_mixed(dn, eng, None)) / (2 * h)
        assert abs(fd - sc[i]) < 1e-6 * max(abs(sc[i]), 1e3), (i, fd, sc[i])


def test_modes_are_stationary(full):
    d, th, fam, eng = full
    m0, _, _ = eng.get_modes()
    st = eng.find_modes(th["betas"], np.linalg.inv(th["D"]), th["phis"], None, warm=True)
    m1, _, _ = eng.get_modes()
    assert st["n_not_converged"] == 0
    assert np.max(np.abs(m1 - m0)) < 1e-9
