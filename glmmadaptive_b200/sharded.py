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
    """The Engine interface over ``world_size`` ranks.  ``local`` is this rank's
    Engine (built from :func:`shard_arrays`); evaluations are all-reduced (sum)
    over ``group`` so every rank returns the whole-data result, exactly what a
    single Engine over all clusters returns."""

    def __init__(self, local, dist=None, group=None, device_tensor_factory=None):
        self.local = local
        self.dist = dist
        self.group = group
        self._buf = None
        self._factory = device_tensor_factory
        for k in ("p", "q", "q_y", "q_zi", "p_zi", "n_phis", "npack", "nAGQ", "K", "result_len"):
            setattr(self, k, getattr(local, k))

    def _buffer(self):
        if self._buf is None:
            import torch
            self._buf = torch.zeros(self.local.result_len, dtype=torch.float64,
                                    device=f"cuda:{self.local.device}")
        return self._buf

    def find_modes(self, betas, invD, phis=None, gammas=None, warm=True):
        # no collective needed: modes are per cluster and stay on the owning GPU
        return self.local.find_modes(betas, invD, phis, gammas, warm)

    def eval_enqueue(self, betas, D, phis=None, gammas=None, want_score=True):
        """Local kernels + all-reduce, all asynchronous on the current stream; returns
        the device tensor holding the reduced result vector."""
        buf = self._buffer()
        self.local.eval_device(betas, D, phis, gammas, want_score, buf.data_ptr())
        if self.dist is not None and self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
            self.dist.all_reduce(buf, group=self.group)
        return buf

    def eval(self, betas, D, phis=None, gammas=None, want_score=True):
        buf = self.eval_enqueue(betas, D, phis, gammas, want_score)
        return self.local.unpack(buf.cpu().numpy(), want_score)


def allreduce_results(vec, dist, group=None):
    """Host-side variant used by the CPU (gloo) tests: sums result vectors over ranks."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64).copy())
    dist.all_reduce(t, group=group)
    return t.numpy()
