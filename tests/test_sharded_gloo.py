"""The N > 1 path on the CPU: world_size-2 gloo processes shard the clusters
(glmmadaptive_b200.sharded), each evaluates its shard (here with the oracle standing in
for the per-GPU engine, which needs CUDA) and the result vectors are all-reduced.  The
sum must equal the unsharded evaluation: log-likelihood and score are sums over
clusters (R/Fit_Funs.R:33-37,105,194,231,271-278)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pack(st, p, n_phis, p_zi, q):
    v = np.zeros(4 + p + n_phis + p_zi + q * q)
    v[0], v[1], v[2] = st["loglik"], st["sum_w"], st["n_nonfinite"]
    v[4:4 + p] = st["g_beta"]
    if n_phis:
        v[4 + p:4 + p + n_phis] = st["g_phi"]
    v[4 + p + n_phis + p_zi:] = st["S"].reshape(-1, order="F")
    return v


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from glmmadaptive_b200 import synth, sharded
    from oracle import agq
    from helpers import oracle_family, oracle_grid, thetas_vec
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, th = synth.make("C3", n=60, ni=6, seed=21, ragged=True)
    fam_name = d.pop("family")
    fam = oracle_family(fam_name)
    arrays = {k: d.get(k) for k in ("y", "X", "Z")}
    loc, (c0, c1), (r0, r1) = sharded.shard_arrays(arrays, d["id"], rank, world)
    od, gh = oracle_grid(loc, th, fam)
    st = agq.suff_stats(thetas_vec(th), od, fam, gh)
    vec = _pack(st, 3, 1, 0, 2)
    tot = sharded.allreduce_results(vec, dist)
    if rank == 0:
        od_all, gh_all = oracle_grid(d, th, fam)
        ref = _pack(agq.suff_stats(thetas_vec(th), od_all, fam, gh_all), 3, 1, 0, 2)
        np.save(out, np.stack([tot, ref]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sum_equals_unsharded(tmp_path):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    tot, ref = np.load(out)
    assert np.allclose(tot, ref, rtol=1e-12, atol=1e-12)
    assert tot[1] == 60.0
