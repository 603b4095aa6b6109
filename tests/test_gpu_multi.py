"""GPU (needs >= 2 devices; skipped on a single-GPU box): the single-process multi-GPU handle
(glmma_create_multi -- what an R session, one process, would use) against the single-device engine
on the same data: every entry point returns the same numbers (sums over devices to 1e-13, concatenated
per-cluster / per-observation outputs exactly), and a whole mixed_model() fit lands on the same
estimates."""
import numpy as np
import pytest

import glmmadaptive_b200 as gb
from glmmadaptive_b200 import synth

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("name", ["C3", "C4"])
def test_multi_device_handle_matches_single_device(name):
    if _ndev() < 2:
        pytest.skip("needs at least two GPUs")
    devs = list(range(min(_ndev(), 4)))
    d, th = synth.make(name, n=3001, ni=7, seed=5, ragged=True)
    fam = d.pop("family")
    rng = np.random.default_rng(2)
    w = rng.uniform(0.5, 1.5, 3001)
    kw = dict(nAGQ=th["nAGQ"], X_zi=d.get("X_zi"), Z_zi=d.get("Z_zi"), weights=w)
    one = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam, device=0, **kw)
    many = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam, device=devs, **kw)
    assert many.n_devices == len(devs) and many.n == one.n == 3001 and many.N == one.N
    betas, D, phis, gammas = th["betas"], th["D"], th.get("phis"), th.get("gammas")
    invD = np.linalg.inv(D)
    s1 = one.find_modes(betas, invD, phis, gammas, warm=False)
    s2 = many.find_modes(betas, invD, phis, gammas, warm=False)
    assert s1["n_not_converged"] == s2["n_not_converged"] == 0 and abs(s1["mean_iterations"] - s2["mean_iterations"]) < 1e-12
    for a, b in zip(one.get_modes(), many.get_modes()):
        assert np.array_equal(a, b)
    r1, r2 = one.eval_raw(betas, D, phis, gammas), many.eval_raw(betas, D, phis, gammas)
    assert np.max(np.abs(r1 - r2) / np.maximum(np.abs(r1), 1.0)) < 1e-13
    e1, e2 = one.em_estep(betas, D, phis, gammas), many.em_estep(betas, D, phis, gammas)
    assert abs(e1["loglik"] - e2["loglik"]) < 1e-13 * abs(e1["loglik"])
    b2 = betas + 0.03
    m1, m2 = one.em_score(b2, phis, gammas), many.em_score(b2, phis, gammas)
    for k in ("g_beta", "g_phi", "g_gamma"):
        if len(m1[k]):
            assert np.max(np.abs(m1[k] - m2[k])) < 1e-11 * max(1.0, np.max(np.abs(m1[k])))
    c1, c2 = one.eval_contrib(betas, D, phis, gammas), many.eval_contrib(betas, D, phis, gammas)
    for k in ("ll_i", "zbar", "S_i"):
        assert np.max(np.abs(c1[k] - c2[k])) <= 1e-13 * max(1.0, np.max(np.abs(c1[k]))), k
    g1, g2 = gb.starting_values_device(one), gb.starting_values_device(many)
    for k in g1:
        assert np.max(np.abs(np.asarray(g1[k]) - np.asarray(g2[k]))) < 1e-10
    # set_modes round trip on the multi handle
    modes, chol, _ = one.get_modes()
    many.set_modes(modes, chol)
    assert np.max(np.abs(many.eval_raw(betas, D, phis, gammas) - r1) / np.maximum(np.abs(r1), 1.0)) < 1e-13
    one.close(); many.close()


def test_multi_device_fit_equals_single_device_fit():
    if _ndev() < 2:
        pytest.skip("needs at least two GPUs")
    d, th = synth.make("C3", n=4000, ni=6, seed=9, ragged=True)
    fam = d.pop("family")
    con = dict(iter_EM=10, iter_qN_outer=3)
    f1 = gb.mixed_model(d["y"], d["X"], d["Z"], d["id"], fam, control=con, device=0, hessian=False)
    f2 = gb.mixed_model(d["y"], d["X"], d["Z"], d["id"], fam, control=con, device=list(range(min(_ndev(), 4))), hessian=False)
    assert f1["converged"] == f2["converged"]
    assert np.max(np.abs(f1["thetas"] - f2["thetas"])) < 1e-8
    assert abs(f1["logLik"] - f2["logLik"]) < 1e-10 * abs(f1["logLik"])
