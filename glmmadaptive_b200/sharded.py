"""Cluster-sharded evaluation across the GPUs of one box: one process per GPU
(torch.distributed), contiguous cluster ranges, and one all-reduce of the small
result vector per evaluation.  Clusters are
conditionally independent, so the log-likelihood and every score component are
plain sums over shards (R/Fit_Funs.R:33-37,105,194,231,244-247,271-278); no data
moves between ranks.
"""
import numpy as np


def shard_bounds(row_ptr, world_size):
    """Contiguous cluster ranges balanced on observations (= on n_i * K point
    evaluations, K being the same for every cluster).  Returns world_size + 1
    cluster indices."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n = len(row_ptr) - 1
    N = row_ptr[-1]
    targets = (np.arange(1, world_size) * N) / world_size
    cuts = np.searchsorted(row_ptr, targets, side="left")
    cuts = np.clip(cuts, 1, n - 1) if n > 1 else cuts
    b = np.concatenate(([0], cuts, [n])).astype(np.int64)
    # keep the ranges non-decreasing even for tiny inputs
    return np.maximum.accumulate(b)


def row_ptr_from_id(ids):
    ids = np.asarray(ids)
    change = np.flatnonzero(ids[1:] != ids[:-1]) + 1
    return np.concatenate(([0], change, [len(ids)])).astype(np.int64)


def shard_arrays(arrays, ids, rank, world_size):
    """Row-slices every per-observation array (and the per-cluster weights) for ``rank``.
    ``arrays`` maps name -> array or None; 'weights' is per cluster, the rest per row."""
    rp = row_ptr_from_id(ids)
    b = shard_bounds(rp, world_size)
    c0, c1 = int(b[rank]), int(b[rank + 1])
    r0, r1 = int(rp[c0]), int(rp[c1])
    out = {}
    for k, v in arrays.items():
        if v is None:
            out[k] = None
        elif k == "weights":
            out[k] = np.asarray(v)[c0:c1]
        else:
            out[k] = np.asarray(v)[r0:r1]
    out["id"] = np.asarray(ids)[r0:r1]
    return out, (c0, c1), (r0, r1)


class ShardedEngine:
    """The Engine interface over ``world_size`` ranks.  ``local`` is this rank's Engine (built
    from :func:`shard_arrays`); every evaluation (eval, E-step, frozen-weight M-step scores) is
    all-reduced (sum) over ``group`` so that each rank holds the whole-data result, exactly what a
    single Engine over all clusters returns.  GHfun / logLik_mixed / score_mixed / mixed_fit take
    a ShardedEngine wherever they take an Engine: every rank runs the same (deterministic) host
    loop on identical reduced numbers, so no parameter broadcast is needed.  Modes stay on the
    owning GPU; ``n`` is the global number of clusters (it enters the D update and score.D)."""

    _SIZES = ("p", "q", "q_y", "q_zi", "p_zi", "n_phis", "npack", "nAGQ", "K", "result_len", "spec", "device")

    def __init__(self, local, dist=None, group=None, reduce_on_host=False, in_engine=True):
        self.local = local
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size(group) > 1) else None
        self.group = group
        self.reduce_on_host = reduce_on_host      # gloo (CPU tests): reduce host vectors
        self._buf = None
        # All-reduce inside the engine (include/glmma.h, glmma_comm_*): the handles' exchange buffers are
        # gathered once through the process group; from then on an evaluation's finalising block sums
        # the ranks' vectors itself over peer memory -- no collective is launched per evaluation.
        self.in_engine = False
        if self.dist is not None and in_engine and not reduce_on_host and hasattr(local, "comm_export") \
                and self.dist.get_world_size(group) <= 16:
            world, rank = self.dist.get_world_size(group), self.dist.get_rank(group)
            handles = [None] * world
            self.dist.all_gather_object(handles, local.comm_export(), group=group)
            local.comm_connect(rank, world, handles)
            self.dist.barrier(group=group)
            self.in_engine = True
        for k in self._SIZES:
            setattr(self, k, getattr(local, k))
        self.n_local = local.n
        self.n = int(self._reduce_host(np.array([float(local.n)]))[0])
        self.N = int(self._reduce_host(np.array([float(local.N)]))[0])
        self._cache_key, self._cache_val = None, None

    # grid bookkeeping lives on the local engine (GHfun / _ensure_grid read and write it there)
    @property
    def _modes_version(self):
        return self.local._modes_version

    @property
    def launch_count(self):
        return self.local.launch_count

    def _reduce_host(self, vec):
        if self.dist is None:
            return np.asarray(vec, float)
        import torch
        t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64).copy())
        if not self.reduce_on_host:
            t = t.to(f"cuda:{self.local.device}")
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def _buffer(self):
        if self._buf is None:
            import torch
            self._buf = torch.zeros(self.local.result_len, dtype=torch.float64,
                                    device=f"cuda:{self.local.device}")
        return self._buf

    def find_modes(self, betas, invD, phis=None, gammas=None, warm=True):
        # no data-path collective: modes are per cluster and stay on the owning GPU; only the
        # convergence counters are summed
        st = self.local.find_modes(betas, invD, phis, gammas, warm)
        self._cache_key = None
        tot = self._reduce_host(np.array([st["n_not_converged"], st["n_chol_failed"],
                                          st["mean_iterations"] * self.n_local], float))
        return dict(n_not_converged=int(tot[0]), n_chol_failed=int(tot[1]),
                    max_iterations=st["max_iterations"], mean_iterations=tot[2] / max(self.n, 1))

    def get_modes(self):
        return self.local.get_modes()

    def set_modes(self, post_modes, chol_upper):
        self.local.set_modes(post_modes, chol_upper)
        self._cache_key = None

    def unpack(self, r, want_score=True):
        return self.local.unpack(r, want_score)

    def eval_enqueue(self, betas, D, phis=None, gammas=None, want_score=True):
        """Local kernels + all-reduce, all asynchronous on the current stream; returns
        the device tensor holding the reduced result vector."""
        buf = self._buffer()
        self.local.eval_device(betas, D, phis, gammas, want_score, buf.data_ptr())
        if self.dist is not None and not self.in_engine:
            self.dist.all_reduce(buf, group=self.group)
        return buf

    def eval(self, betas, D, phis=None, gammas=None, want_score=True):
        if self.reduce_on_host:
            r = self._reduce_host(self.local.eval_raw(betas, D, phis, gammas, want_score))
            return self.local.unpack(r, want_score)
        if self.dist is None or self.in_engine:
            # the plain C-ABI call (glmma_eval: host theta in, host result vector out); with the handles
            # connected the vector it returns is already the sum over the ranks
            return self.local.unpack(self.local.eval_raw(betas, D, phis, gammas, want_score), want_score)
        buf = self.eval_enqueue(betas, D, phis, gammas, want_score)
        return self.local.unpack(buf.cpu().numpy(), want_score)

    def em_estep(self, betas, D, phis=None, gammas=None):
        r = self.local.em_estep_raw(betas, D, phis, gammas)
        return self.local.unpack(r if self.in_engine else self._reduce_host(r), True)

    def em_score(self, betas, phis=None, gammas=None):
        r = self.local.em_score_raw(betas, phis, gammas)
        return self.local.unpack(r if self.in_engine else self._reduce_host(r), True)

    def observed_information(self, *a, **k):
        H = self.local.observed_information(*a, **k)
        return self._reduce_host(H.reshape(-1)).reshape(H.shape)

    def glm_pass(self, *a, **k):
        return self._reduce_host(self.local.glm_pass(*a, **k))

    def eval_contrib(self, *a, **k):
        return self.local.eval_contrib(*a, **k)      # per-cluster / per-observation pieces stay local

    def close(self):
        self.local.close()


def allreduce_results(vec, dist, group=None):
    """Host-side variant used by the CPU (gloo) tests: sums result vectors over ranks."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64).copy())
    dist.all_reduce(t, group=group)
    return t.numpy()
