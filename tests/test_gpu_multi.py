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
    # different shards sum in a different order: 1e-13 of the LARGEST component bounds the rounding of every sum (the
    # phi-score is the small difference of two sums of the size of the log-likelihood: node-independent terms added
    # once per evaluation against the per-node terms), 1e-11 relative what is left after that cancellation
    assert np.max(np.abs(r1 - r2)) < 1e-13 * np.max(np.abs(r1))
    assert np.max(np.abs(r1 - r2) / np.maximum(np.abs(r1), 1.0)) < 1e-11
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
    assert np.max(np.abs(many.eval_raw(betas, D, phis, gammas) - r1)) < 1e-13 * np.max(np.abs(r1))
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


# ---- one process per GPU: the all-reduce inside the engine (glmma_comm_export / glmma_comm_connect) ----
def _comm_worker(rank, world, port, out_dir):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    from glmmadaptive_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, th = synth.make("C3", n=4001, ni=7, seed=9, ragged=True)
    fam = d.pop("family")
    arrays = {k: d.get(k) for k in ("y", "X", "Z")}
    loc, _, _ = sharded.shard_arrays(arrays, d["id"], rank, world)
    eng = gb.Engine(loc["y"], loc["X"], loc["Z"], loc["id"], fam, nAGQ=th["nAGQ"], device=rank)
    sh = sharded.ShardedEngine(eng, dist)
    assert sh.in_engine and eng.comm_size == world
    betas, D, phis = th["betas"], th["D"], th["phis"]
    sh.find_modes(betas, np.linalg.inv(D), phis, None, warm=False)
    outs = []
    for k in range(6):                      # several evaluations: both parities of the exchange slots
        r = sh.eval(betas * (1.0 + 0.01 * k), D, phis, None)
        outs.append(np.concatenate([[r["loglik"], r["sum_w"], r["n_nonfinite"]], r["g_beta"], r["g_phi"], r["S"].ravel()]))
    es = sh.em_estep(betas, D, phis, None)
    sc = sh.em_score(betas * 1.02, phis, None)
    outs.append(np.concatenate([[es["loglik"], es["sum_w"], es["n_nonfinite"]], es["g_beta"], es["g_phi"], es["S"].ravel()]))
    outs.append(np.concatenate([[0.0, sc["sum_w"], sc["n_nonfinite"]], sc["g_beta"], sc["g_phi"], sc["S"].ravel()]))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.stack(outs))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


def test_in_engine_allreduce_matches_unsharded(tmp_path):
    """Two processes, one GPU each, handles connected through peer memory: every rank gets the same
    bits, equal (1e-11 relative: the shards are summed in a different order, and the phi-score cancels) to the unsharded engine's evaluation of the whole data set."""
    if _ndev() < 2:
        pytest.skip("needs at least two GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_comm_worker, args=(world, 29561, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(str(tmp_path / f"rank{r}.npy")) for r in range(world)]
    assert np.array_equal(res[0], res[1])
    d, th = synth.make("C3", n=4001, ni=7, seed=9, ragged=True)
    fam = d.pop("family")
    one = gb.Engine(d["y"], d["X"], d["Z"], d["id"], fam, nAGQ=th["nAGQ"], device=0)
    betas, D, phis = th["betas"], th["D"], th["phis"]
    one.find_modes(betas, np.linalg.inv(D), phis, None, warm=False)
    for k in range(6):
        r = one.eval(betas * (1.0 + 0.01 * k), D, phis, None)
        ref = np.concatenate([[r["loglik"], r["sum_w"], r["n_nonfinite"]], r["g_beta"], r["g_phi"], r["S"].ravel()])
        assert np.max(np.abs(res[0][k] - ref) / np.maximum(1.0, np.abs(ref))) < 1e-11, k
    es = one.em_estep(betas, D, phis, None)
    ref = np.concatenate([[es["loglik"], es["sum_w"], es["n_nonfinite"]], es["g_beta"], es["g_phi"], es["S"].ravel()])
    assert np.max(np.abs(res[0][6] - ref) / np.maximum(1.0, np.abs(ref))) < 1e-11
    sc = one.em_score(betas * 1.02, phis, None)
    ref = np.concatenate([[0.0, sc["sum_w"], sc["n_nonfinite"]], sc["g_beta"], sc["g_phi"], sc["S"].ravel()])
    assert np.max(np.abs(res[0][7] - ref) / np.maximum(1.0, np.abs(ref))) < 1e-11
    one.close()
