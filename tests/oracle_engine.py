"""TEST INFRASTRUCTURE: a CPU stand-in for glmmadaptive_b200.Engine built on the oracle, so that
the product's HOST logic (glmmadaptive_b200/fit.py: mixed_fit's EM / quasi-Newton loop;
glmmadaptive_b200/sharded.py: the cluster-sharded all-reduce path) can be exercised where there
is no GPU -- single process and world_size-2 gloo.  It answers the same calls as the ctypes
Engine (find_modes / get_modes / set_modes / eval_raw / em_estep_raw / em_score_raw / unpack)
with oracle arithmetic.  Never imported by the product."""
import numpy as np

from oracle import agq, fit as ofit
from glmmadaptive_b200.engine import family_spec
from helpers import oracle_data, oracle_family, pack_chol


class OracleEngine:
    def __init__(self, d, fam_name, nAGQ, link=None):
        self.spec = family_spec(fam_name, link)
        self.fam = oracle_family(fam_name, link)
        self.od = oracle_data(d)
        self.N, self.n = self.od.N, self.od.n
        self.p, self.q_y, self.q_zi = self.od.X.shape[1], self.od.q_y, self.od.q_zi
        self.p_zi = 0 if self.od.X_zi is None else self.od.X_zi.shape[1]
        self.q = self.q_y + self.q_zi
        self.npack = self.q * (self.q + 1) // 2
        self.n_phis = self.spec.n_phis
        self.nAGQ = int(nAGQ) if nAGQ is not None else (11 if self.q < 3 else 7)
        self.K = self.nAGQ ** self.q
        self.result_len = 4 + self.p + self.n_phis + self.p_zi + self.q * self.q
        self.device = 0
        self.launch_count = 0
        self._modes_version = None
        self._cache_key = self._cache_val = None
        self._modes = np.zeros((self.n, self.q))
        self._gh = None
        self._estep = None

    # -- grid -----------------------------------------------------------------
    def find_modes(self, betas, invD, phis=None, gammas=None, warm=True):
        start = self._modes if warm else np.zeros((self.n, self.q))
        modes, hess, iters = agq.find_modes_newton(start, self.od, self.fam, np.asarray(betas, float),
                                                   np.asarray(invD, float), phis, gammas)
        self._modes = modes
        self._gh = agq.GHfun_from_modes(self.od, modes, hess, self.nAGQ, self.q)
        self._modes_version = object()
        self._cache_key = None
        return dict(n_not_converged=0, n_chol_failed=0, max_iterations=int(np.max(iters)),
                    mean_iterations=float(np.mean(iters)))

    def get_modes(self):
        return self._gh.post_modes.copy(), pack_chol(self._gh.chol_hessians), self._gh.log_dets.copy()

    def set_modes(self, post_modes, chol_upper):
        q = self.q
        idx = [(r, c) for c in range(q) for r in range(c + 1)]
        hess = []
        for row in np.asarray(chol_upper, float).reshape(self.n, self.npack):
            R = np.zeros((q, q))
            for v, (r, c) in zip(row, idx):
                R[r, c] = v
            hess.append(R.T @ R)
        self._modes = np.asarray(post_modes, float).reshape(self.n, q).copy()
        self._gh = agq.GHfun_from_modes(self.od, self._modes, hess, self.nAGQ, q)
        self._modes_version = object()
        self._cache_key = None

    # -- evaluations ------------------------------------------------------------
    def _pack(self, st):
        v = np.zeros(self.result_len)
        v[0], v[1], v[2] = st["loglik"], st["sum_w"], st["n_nonfinite"]
        o = 4
        v[o:o + self.p] = st["g_beta"]; o += self.p
        if self.n_phis:
            v[o:o + self.n_phis] = st["g_phi"]; o += self.n_phis
        if self.p_zi:
            v[o:o + self.p_zi] = st["g_gamma"]; o += self.p_zi
        v[o:] = st["S"].reshape(-1, order="F")
        return v

    def _thetas(self, betas, D, phis, gammas):
        parts = [np.asarray(betas, float), agq.chol_transf(np.asarray(D, float).reshape(self.q, self.q))]
        if self.n_phis:
            parts.append(np.atleast_1d(phis))
        if self.p_zi:
            parts.append(np.atleast_1d(gammas))
        return np.concatenate(parts)

    def eval_raw(self, betas, D, phis=None, gammas=None, want_score=True):
        return self._pack(agq.suff_stats(self._thetas(betas, D, phis, gammas), self.od, self.fam, self._gh))

    def unpack(self, r, want_score=True):
        out = dict(loglik=float(r[0]), sum_w=float(r[1]), n_nonfinite=int(r[2]))
        o = 4
        out["g_beta"] = np.array(r[o:o + self.p]); o += self.p
        out["g_phi"] = np.array(r[o:o + self.n_phis]); o += self.n_phis
        out["g_gamma"] = np.array(r[o:o + self.p_zi]); o += self.p_zi
        out["S"] = np.array(r[o:o + self.q * self.q]).reshape(self.q, self.q)
        return out

    def eval(self, betas, D, phis=None, gammas=None, want_score=True):
        return self.unpack(self.eval_raw(betas, D, phis, gammas, want_score))

    def em_estep_raw(self, betas, D, phis=None, gammas=None):
        D = np.asarray(D, float).reshape(self.q, self.q)
        p_by, post_b2, ll, eta_zi = ofit.e_step(self.od, self.fam, self._gh, np.asarray(betas, float), D, phis, gammas)
        self._estep = (p_by, eta_zi)
        return self.eval_raw(betas, D, phis, gammas)

    def em_score_raw(self, betas, phis=None, gammas=None):
        p_by, _ = self._estep
        betas = np.asarray(betas, float)
        eta_zi = None if self.p_zi == 0 else ofit._etas(self.od, self._gh, betas, np.asarray(gammas, float))[1]
        v = np.zeros(self.result_len)
        o = 4
        v[o:o + self.p] = -ofit.score_betas(betas, self.od, self.fam, self._gh, phis, eta_zi, p_by); o += self.p
        if self.n_phis:
            v[o:o + self.n_phis] = -ofit.score_phis(np.atleast_1d(phis), self.od, self.fam, self._gh, betas, eta_zi, p_by)
            o += self.n_phis
        if self.p_zi:
            v[o:o + self.p_zi] = -ofit.score_gammas(np.asarray(gammas, float), self.od, self.fam, self._gh, betas, phis, p_by)
        return v

    def em_estep(self, betas, D, phis=None, gammas=None):
        return self.unpack(self.em_estep_raw(betas, D, phis, gammas))

    def em_score(self, betas, phis=None, gammas=None):
        return self.unpack(self.em_score_raw(betas, phis, gammas))

    def close(self):
        pass
