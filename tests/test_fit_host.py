"""CPU: the product's host-side fitting loop (glmmadaptive_b200/fit.py) and its cluster-sharded
form (glmmadaptive_b200/sharded.py, world_size 2 over gloo) with the oracle standing in for the
per-GPU engine (tests/oracle_engine.py).  The loop must land on the same estimates as the
oracle's independent restatement of R/mixed_fit.R (oracle/fit.py), and the sharded fit on the
same estimates as the unsharded one."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import fit as ofit
import glmmadaptive_b200 as gb
from glmmadaptive_b200 import synth
from helpers import oracle_data, oracle_family
from oracle_engine import OracleEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CON = dict(nAGQ=7, iter_EM=8, iter_qN_outer=3)


def _data(name="C3", n=40):
    d, th = synth.make(name, n=n, ni=5, seed=33, ragged=True)
    return d, d.pop("family")


@pytest.mark.parametrize("name", ["C3", "C4"])
def test_host_loop_matches_oracle_fit(name):
    d, fam_name = _data(name)
    eng = OracleEngine(d, fam_name, CON["nAGQ"])
    inits = gb.starting_values(d["y"], d["X"], d.get("X_zi"), eng.spec.family, eng.q, eng.n_phis)
    fit = gb.mixed_fit(eng, inits, CON, hessian=True)
    ref = ofit.mixed_fit(oracle_data(d), oracle_family(fam_name), control=CON, mode_finder="newton", hessian=True)
    assert fit["converged"] == ref["converged"] and fit["n_outer"] == ref["n_outer"]
    assert np.max(np.abs(fit["thetas"] - ref["thetas"])) < 1e-8
    assert abs(fit["logLik"] - ref["logLik"]) < 1e-9 * abs(ref["logLik"])
    assert np.max(np.abs(fit["Hessian"] - ref["Hessian"])) < 1e-6 * np.max(np.abs(ref["Hessian"]))
    assert fit["counts"]["estep"] >= 5 and fit["counts"]["em_score"] > 0


def test_host_loop_penalized_matches_oracle_fit():
    """penalized = list(pen_mu, pen_sigma, pen_df): the Student's t penalty on the fixed effects in the
    log-likelihood, the score, the EM log-likelihood and the M-step beta score (R/Fit_Funs.R:39-40,
    169-173, 360-364; R/mixed_fit.R:156-159) -- product host code against the oracle's restatement."""
    d, fam_name = _data("C3")
    pen = dict(pen_mu=0.1, pen_sigma=0.5, pen_df=4.0)
    eng = OracleEngine(d, fam_name, CON["nAGQ"])
    inits = gb.starting_values(d["y"], d["X"], d.get("X_zi"), eng.spec.family, eng.q, eng.n_phis)
    fit = gb.mixed_fit(eng, inits, CON, hessian=True, penalized=pen)
    ref = ofit.mixed_fit(oracle_data(d), oracle_family(fam_name), control=CON, mode_finder="newton", hessian=True,
                         penalized=pen)
    plain = ofit.mixed_fit(oracle_data(d), oracle_family(fam_name), control=CON, mode_finder="newton", hessian=False)
    assert np.max(np.abs(fit["thetas"] - ref["thetas"])) < 1e-8
    assert abs(fit["logLik"] - ref["logLik"]) < 1e-9 * abs(ref["logLik"])
    assert np.max(np.abs(fit["Hessian"] - ref["Hessian"])) < 1e-6 * np.max(np.abs(ref["Hessian"]))
    assert np.max(np.abs(ref["thetas"] - plain["thetas"])) > 1e-3          # the penalty does something
    with pytest.raises(ValueError):
        gb.Penalty.from_arg(dict(sigma=1.0), 3)


def test_glm_starts_follow_the_binomial_link():
    """binomial(link = "probit" / "cloglog"): glm.fit runs with the model's own link
    (R/mixed_model.R:176); the IRLS solution satisfies that link's score equations."""
    from scipy.special import ndtr
    from glmmadaptive_b200 import fit as hf
    rng = np.random.default_rng(4)
    X = np.column_stack([np.ones(400), rng.normal(size=400)])
    eta = X @ np.array([-0.3, 0.7])
    for link, inv, dmu in (("probit", ndtr, lambda e: np.exp(-0.5 * e * e) / np.sqrt(2 * np.pi)),
                           ("cloglog", lambda e: -np.expm1(-np.exp(e)), lambda e: np.exp(e - np.exp(e)))):
        y = (rng.random(400) < inv(eta)).astype(float)
        b = hf.glm_fit(X, y, "binomial", link=link)
        e = X @ b
        mu = inv(e)
        sc = X.T @ ((y - mu) * dmu(e) / (mu * (1 - mu)))
        assert np.max(np.abs(sc)) < 1e-3, (link, sc)       # glm.fit stops on the deviance (1e-8 relative)
        assert np.max(np.abs(b - hf.glm_fit(X, y, "binomial"))) > 1e-2      # not the logit fit


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from glmmadaptive_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, fam_name = _data("C3", 40)
    arrays = {k: d.get(k) for k in ("y", "X", "Z")}
    loc, _, _ = sharded.shard_arrays(arrays, d["id"], rank, world)
    eng = sharded.ShardedEngine(OracleEngine(loc, fam_name, CON["nAGQ"]), dist, reduce_on_host=True)
    assert eng.n == 40 and eng.n_local < 40
    # starting values come from the whole data set on every rank (host-side glm.fit, R/mixed_model.R:162)
    inits = gb.starting_values(d["y"], d["X"], None, eng.spec.family, eng.q, eng.n_phis)
    fit = gb.mixed_fit(eng, inits, CON, hessian=False)
    allt = [None] * world
    dist.all_gather_object(allt, fit["thetas"])
    if rank == 0:
        np.save(out, np.stack(allt + [np.full_like(fit["thetas"], fit["logLik"])]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_fit_equals_unsharded(tmp_path):
    out = str(tmp_path / "fit.npy")
    mp.spawn(_worker, args=(2, 29547, out), nprocs=2, join=True)
    res = np.load(out)
    assert np.array_equal(res[0], res[1])            # every rank walks the same path
    d, fam_name = _data("C3", 40)
    eng = OracleEngine(d, fam_name, CON["nAGQ"])
    inits = gb.starting_values(d["y"], d["X"], None, eng.spec.family, eng.q, eng.n_phis)
    one = gb.mixed_fit(eng, inits, CON, hessian=False)
    assert np.max(np.abs(res[0] - one["thetas"])) < 1e-7
    assert abs(res[2][0] - one["logLik"]) < 1e-8 * abs(one["logLik"])
