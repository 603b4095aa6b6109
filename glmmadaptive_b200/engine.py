"""Host-side mirror of the reference's hot-path closures over the CUDA engine.

Reference seam (paths relative to the reference tree):

  GHfun()        R/Functions.R:105-138   -> :func:`GHfun`
  logLik_mixed() R/Fit_Funs.R:1-42       -> :func:`logLik_mixed`
  score_mixed()  R/Fit_Funs.R:44-293     -> :func:`score_mixed`

The argument meaning and the returned shapes follow the reference; the lists of
per-cluster data and the family closures that the R functions receive are
replaced by one :class:`Engine` (the device-resident data set + family id).
All arithmetic on N- or n-sized data happens on the GPU through libglmma.so;
this module only does the theta <-> (betas, D, phis, gammas) bookkeeping and the
q x q chain rule for the D parameters (R/Fit_Funs.R:237-286).
There is no CPU fallback.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import GlmmaData, GlmmaError, GlmmaModeStats, FAMILY_IDS, LINK_IDS

# reference constructor / family$family  ->  (family$family, default link, n_phis, zero part, y columns)
_SPECS = {
    "binomial": ("binomial", "logit", 0, False, None),
    "poisson": ("poisson", "log", 0, False, 1),
    "negative.binomial": ("negative binomial", "log", 1, False, 1),
    "zi.poisson": ("zero-inflated poisson", "log", 0, True, 1),
    "zi.negative.binomial": ("zero-inflated negative binomial", "log", 1, True, 1),
    "zi.binomial": ("zero-inflated binomial", "logit", 0, True, None),
    "hurdle.poisson": ("hurdle poisson", "log", 0, True, 1),
    "hurdle.negative.binomial": ("hurdle negative binomial", "log", 1, True, 1),
    "hurdle.lognormal": ("hurdle log-normal", "identity", 1, True, 1),
    "beta.fam": ("beta", "logit", 1, False, 1),
    "hurdle.beta.fam": ("hurdle beta", "logit", 1, True, 1),
    "Gamma.fam": ("Gamma", "log", 1, False, 1),
    "censored.normal": ("censored normal", "identity", 1, False, 2),
    "students.t": ("Student's-t", "identity", 1, False, 1),
    "beta.binomial": ("beta binomial", "logit", 1, False, None),
}
_BY_FAMILY_STRING = {v[0]: v for v in _SPECS.values()}


@dataclass(frozen=True)
class FamilySpec:
    family: str          # family$family of the reference
    link: str
    n_phis: int
    has_zi: bool


def family_spec(name, link=None):
    """Accepts the reference's constructor name (``"negative.binomial"``) or its
    ``family$family`` string (``"negative binomial"``)."""
    rec = _SPECS.get(name) or _BY_FAMILY_STRING.get(name)
    if rec is None:
        raise ValueError(f"family {name!r} is not accelerated by the engine (it stays on the reference path)")
    fam, dlink, n_phis, zi, _ = rec
    link = dlink if link is None else link
    if fam == "binomial":
        if link not in ("logit", "probit", "cloglog"):
            raise ValueError(f"binomial link {link!r} is not accelerated")
    elif fam == "beta binomial":
        if link not in ("logit", "cloglog"):
            raise ValueError(f"beta.binomial link {link!r} is not accelerated")
    elif link != dlink:
        raise ValueError(f"{fam}: only the {dlink} link is accelerated")
    return FamilySpec(fam, link, n_phis, zi)


def _dp(a):
    return None if a is None else a.ctypes.data_as(_lib.c_double_p)


def _f(a, order="F"):
    return None if a is None else np.require(np.asarray(a, dtype=np.float64), requirements=["A", "O"] + [order])


class Engine:
    """Device-resident data set + family: what ``mixed_fit()`` receives
    (R/mixed_fit.R:1-25), uploaded once.  Wraps one ``glmma_handle``."""

    def __init__(self, y, X, Z, id, family, link=None, nAGQ=None, X_zi=None, Z_zi=None, offset=None,
                 offset_zi=None, weights=None, device=0, df=None):
        """``device``: a CUDA device index, or a list of indices to spread the clusters over several
        GPUs of this process (glmma_create_multi)."""
        self.lib = _lib.load()
        self.spec = family_spec(family, link) if not isinstance(family, FamilySpec) else family
        y = _f(y)
        X = _f(X)
        Z = _f(np.asarray(Z, float).reshape(len(X), -1))
        X_zi = None if X_zi is None else _f(np.asarray(X_zi, float).reshape(len(X), -1))
        Z_zi = None if Z_zi is None else _f(np.asarray(Z_zi, float).reshape(len(X), -1))
        offset, offset_zi, weights = _f(offset), _f(offset_zi), _f(weights)
        # Cluster index: rows of one cluster must be contiguous (the reference sorts its data frame by
        # the grouping factor, R/mixed_model.R:50).  Any label type works: only label changes matter.
        labels = np.asarray(id).reshape(-1)
        if len(labels) != X.shape[0]:
            raise ValueError("id must have one entry per row of X")
        change = np.flatnonzero(labels[1:] != labels[:-1]) + 1
        row_ptr = np.ascontiguousarray(np.concatenate(([0], change, [len(labels)])), dtype=np.int64)
        # a label that heads two runs means the rows are not grouped: checked on the run heads only
        # (n values, not N), for every size
        heads = labels[row_ptr[:-1]]
        if len(np.unique(heads)) != len(heads):
            raise ValueError("rows are not grouped by id: sort the data by the grouping factor first")
        N = X.shape[0]
        self.N, self.p, self.q_y = N, X.shape[1], Z.shape[1]
        self.p_zi = 0 if X_zi is None else X_zi.shape[1]
        self.q_zi = 0 if Z_zi is None else Z_zi.shape[1]
        self.q = self.q_y + self.q_zi
        # control$nAGQ default, R/mixed_model.R:154
        self.nAGQ = int(nAGQ) if nAGQ is not None else (11 if self.q < 3 else 7)
        self.n_phis = self.spec.n_phis
        d = GlmmaData()
        d.family = FAMILY_IDS[self.spec.family]
        d.link = LINK_IDS[self.spec.link]
        d.n_obs, d.n_clusters = N, 0
        d.p, d.q_y, d.p_zi, d.q_zi = self.p, self.q_y, self.p_zi, self.q_zi
        d.n_phis, d.y_cols, d.nAGQ = self.n_phis, (2 if y.ndim == 2 and y.shape[1] == 2 else 1), self.nAGQ
        d.y, d.X, d.Z, d.X_zi, d.Z_zi = _dp(y), _dp(X), _dp(Z), _dp(X_zi), _dp(Z_zi)
        d.offset, d.offset_zi, d.weights = _dp(offset), _dp(offset_zi), _dp(weights)
        d.id = None
        d.row_ptr = row_ptr.ctypes.data_as(C.POINTER(C.c_int64))
        d.n_clusters = len(row_ptr) - 1
        d.family_df = 0.0 if df is None else float(df)      # students.t(df)
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            # one process, several GPUs: contiguous cluster ranges per device behind one handle
            devs = (C.c_int * len(device))(*[int(v) for v in device])
            rc = self.lib.glmma_create_multi(C.byref(d), devs, len(device), C.byref(h))
            device = device[0]
        else:
            rc = self.lib.glmma_create(C.byref(d), int(device), C.byref(h))
        if rc != 0:
            raise GlmmaError(rc, self.lib.glmma_last_error(None).decode())
        self._h = h
        self._stage, self._res, self._res_p = None, None, None
        self.device = int(device)
        self.n_devices = int(self.lib.glmma_n_devices(h))
        self.n = int(self.lib.glmma_n_clusters(h))
        self.K = int(self.lib.glmma_n_nodes(h))
        self.result_len = int(self.lib.glmma_result_len(h))
        self.npack = self.q * (self.q + 1) // 2
        self._modes_version = None
        self._cache_key, self._cache_val = None, None
        self._X_host, self._X_zi_host = X, X_zi
        self.input_bytes = sum(a.nbytes for a in (y, X, Z, X_zi, Z_zi, offset, offset_zi, weights) if a is not None) + \
            4 * len(row_ptr)

    # -- lifetime -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.glmma_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise GlmmaError(rc, self.lib.glmma_last_error(self._h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.glmma_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    @property
    def launch_count(self):
        return int(self.lib.glmma_launch_count(self._h))

    # -- GHfun pieces ---------------------------------------------------------
    def _theta_ptrs(self, betas, M, phis, gammas):
        """theta into the engine's persistent staging buffer (betas | M column-major | phis | gammas) and
        the four pointers into it: an optimiser calls this thousands of times, so nothing is allocated
        and no ctypes object is created per call."""
        st = self._stage
        if st is None:
            p, q, nph, pz = self.p, self.q, self.n_phis, self.p_zi
            buf = np.zeros(p + q * q + nph + pz)
            base = buf.ctypes.data
            ptr = lambda off, n: C.cast(base + 8 * off, _lib.c_double_p) if n else None
            st = self._stage = dict(buf=buf, b=buf[:p], M=buf[p:p + q * q].reshape(q, q, order="F"),
                                    ph=buf[p + q * q:p + q * q + nph], g=buf[p + q * q + nph:],
                                    ptrs=(ptr(0, p), ptr(p, q * q), ptr(p + q * q, nph), ptr(p + q * q + nph, pz)))
        if np.size(betas) != self.p:
            raise ValueError("betas has the wrong length")
        st["b"][:] = betas
        st["M"][...] = np.reshape(M, (self.q, self.q))
        if self.n_phis:
            if phis is None or np.size(phis) != self.n_phis:
                raise ValueError("phis has the wrong length")
            st["ph"][:] = phis
        if self.p_zi:
            if gammas is None or np.size(gammas) != self.p_zi:
                raise ValueError("gammas has the wrong length")
            st["g"][:] = gammas
        return st["buf"], st["ptrs"]

    def find_modes(self, betas, invD, phis=None, gammas=None, warm=True):
        keep, ptrs = self._theta_ptrs(betas, invD, phis, gammas)
        st = GlmmaModeStats()
        self._check(self.lib.glmma_find_modes(self._h, *ptrs, int(bool(warm)), C.byref(st)))
        self._modes_version = object()
        self._cache_key = None
        return dict(n_not_converged=st.n_not_converged, n_chol_failed=st.n_chol_failed,
                    max_iterations=st.max_iterations, mean_iterations=st.mean_iterations)

    def get_modes(self):
        modes = np.empty((self.n, self.q), order="F")
        chol = np.empty((self.n, self.npack))
        logdet = np.empty(self.n)
        self._check(self.lib.glmma_get_modes(self._h, _dp(modes), _dp(chol), _dp(logdet)))
        return modes, chol, logdet

    def set_modes(self, post_modes, chol_upper):
        m = _f(np.asarray(post_modes, float).reshape(self.n, self.q))
        c = _f(np.asarray(chol_upper, float).reshape(self.n, self.npack), "C")
        self._check(self.lib.glmma_set_modes(self._h, _dp(m), _dp(c)))
        self._modes_version = object()
        self._cache_key = None

    # -- evaluation -----------------------------------------------------------
    def unpack(self, r, want_score=True):
        out = dict(loglik=float(r[0]), sum_w=float(r[1]), n_nonfinite=int(r[2]))
        if want_score:
            o = 4
            out["g_beta"] = np.array(r[o:o + self.p]); o += self.p
            out["g_phi"] = np.array(r[o:o + self.n_phis]); o += self.n_phis
            out["g_gamma"] = np.array(r[o:o + self.p_zi]); o += self.p_zi
            out["S"] = np.array(r[o:o + self.q * self.q]).reshape(self.q, self.q)
        return out

    def eval_raw(self, betas, D, phis=None, gammas=None, want_score=True):
        """The raw result vector of glmma_eval (layout in include/glmma.h)."""
        keep, ptrs = self._theta_ptrs(betas, D, phis, gammas)
        if self._res is None:
            self._res = np.zeros(self.result_len)
            self._res_p = _dp(self._res)
        self._res[4:] = 0.0
        self._check(self.lib.glmma_eval(self._h, *ptrs, 1 if want_score else 0, self._res_p))
        return self._res.copy()

    def eval(self, betas, D, phis=None, gammas=None, want_score=True):
        return self.unpack(self.eval_raw(betas, D, phis, gammas, want_score), want_score)

    def eval_device(self, betas, D, phis, gammas, want_score, device_ptr):
        """Enqueues one evaluation; the result vector is left at ``device_ptr`` (no sync)."""
        keep, ptrs = self._theta_ptrs(betas, D, phis, gammas)
        self._check(self.lib.glmma_eval_device(self._h, *ptrs, int(bool(want_score)), C.c_void_p(device_ptr)))

    def eval_contrib(self, betas, D, phis=None, gammas=None):
        keep, ptrs = self._theta_ptrs(betas, D, phis, gammas)
        ll_i = np.empty(self.n)
        zbar = np.empty(self.N)
        zbar_zi = np.empty(self.N) if self.p_zi else None
        sphi = np.empty(self.N) if self.n_phis else None
        S_i = np.empty((self.n, self.q, self.q))
        self._check(self.lib.glmma_eval_contrib(self._h, *ptrs, _dp(ll_i), _dp(zbar), _dp(zbar_zi), _dp(sphi), _dp(S_i)))
        return dict(ll_i=ll_i, zbar=zbar, zbar_zi=zbar_zi, sphi_obs=sphi, S_i=S_i)

    # -- EM phase (R/mixed_fit.R:119-229) ---------------------------------------
    def em_estep(self, betas, D, phis=None, gammas=None):
        """E-step on the current grid: glmma_eval's result (loglik = lgLik[it], S = sum_i w_i
        E[b b' | y_i]) and the posterior node weights p_by kept on the device."""
        return self.unpack(self.em_estep_raw(betas, D, phis, gammas), True)

    def em_estep_raw(self, betas, D, phis=None, gammas=None):
        keep, ptrs = self._theta_ptrs(betas, D, phis, gammas)
        res = np.zeros(self.result_len)
        self._check(self.lib.glmma_em_estep(self._h, *ptrs, _dp(res)))
        return res

    def em_score(self, betas, phis=None, gammas=None):
        """Scores at (betas, phis, gammas) with p_by frozen at the last E-step
        (score_betas / score_phis / score_gammas, R/Fit_Funs.R:295-427, up to sign)."""
        return self.unpack(self.em_score_raw(betas, phis, gammas), True)

    def em_score_raw(self, betas, phis=None, gammas=None):
        keep, ptrs = self._theta_ptrs(betas, np.eye(self.q), phis, gammas)
        res = np.zeros(self.result_len)
        self._check(self.lib.glmma_em_score(self._h, ptrs[0], ptrs[2], ptrs[3], _dp(res)))
        return res

    # -- starting values (R/mixed_model.R:157-194) ---------------------------------
    GLM_FAMILIES = {"binomial": 0, "poisson": 1, "Gamma": 2, "gaussian": 3}

    def glm_pass(self, glm_family, design="X", response="y", use_offset=True, beta=None):
        """One IRLS pass of stats::glm.fit on the resident rows: [deviance, X'Wz (p), X'WX (p x p)]."""
        ncol = self.p if design == "X" else self.p_zi
        out = np.zeros(1 + ncol + ncol * ncol)
        b = None if beta is None else _f(np.atleast_1d(beta), "C")
        self._check(self.lib.glmma_glm_pass(self._h, 0 if design == "X" else 1, self.GLM_FAMILIES[glm_family],
                                            0 if response == "y" else 1, int(bool(use_offset)),
                                            int(beta is None), _dp(b), _dp(out)))
        return out

    # -- sharded jobs, one process per GPU: all-reduce inside the engine (include/glmma.h) --------
    def comm_export(self):
        buf = C.create_string_buffer(64)
        self._check(self.lib.glmma_comm_export(self._h, buf))
        return bytes(buf.raw)

    def comm_connect(self, rank, nranks, handles):
        """handles: the nranks exported byte strings in rank order."""
        blob = b"".join(handles)
        assert len(blob) == 64 * nranks
        buf = C.create_string_buffer(blob, len(blob))
        self._check(self.lib.glmma_comm_connect(self._h, int(rank), int(nranks), buf))

    @property
    def comm_size(self):
        return int(self.lib.glmma_comm_size(self._h))

    def observed_information(self, betas, D, phis=None, gammas=None, diag_D=False):
        """Hessian of the NEGATIVE log-likelihood on the current grid w.r.t. c(betas, D-params, phis,
        gammas) from one device pass (glmma_observed_information); what the reference computes as
        cd_vec(score_mixed), R/mixed_fit.R:324-342."""
        nD = self.q if diag_D else self.npack
        P = self.p + nD + self.n_phis + self.p_zi
        keep, ptrs = self._theta_ptrs(betas, D, phis, gammas)
        H = np.zeros((P, P), order="F")
        self._check(self.lib.glmma_observed_information(self._h, *ptrs, int(bool(diag_D)), _dp(H)))
        return np.ascontiguousarray(H)

    def log_post_b(self, b, betas, invD, phis=None, gammas=None):
        """log_post_b of predict()'s Metropolis step (R/methods.R:1020-1036) for one random-effects
        vector per cluster: b is n x q; returns the n values sum_j log f(y_ij | b_i) - b_i' invD b_i / 2."""
        b = _f(np.asarray(b, float).reshape(self.n, self.q))
        keep, ptrs = self._theta_ptrs(betas, invD, phis, gammas)
        out = np.zeros(self.n)
        self._check(self.lib.glmma_log_post_b(self._h, _dp(b), *ptrs, _dp(out)))
        return out

    def time_eval(self, betas, D, phis=None, gammas=None, want_score=True, iters=10):
        keep, ptrs = self._theta_ptrs(betas, D, phis, gammas)
        a, b = C.c_float(), C.c_float()
        self._check(self.lib.glmma_time_eval(self._h, *ptrs, int(bool(want_score)), int(iters), C.byref(a), C.byref(b)))
        return a.value, b.value


def measure_fp64_peak(device=0):
    lib = _lib.load()
    v = C.c_double()
    rc = lib.glmma_measure_fp64_peak(int(device), C.byref(v))
    if rc != 0:
        raise GlmmaError(rc, lib.glmma_last_error(None).decode())
    return v.value


# ----------------------------------------------------------------------------
# theta bookkeeping (R/mixed_fit.R:240-247; R/Functions.R:140-158)
# ----------------------------------------------------------------------------
def _upper_tri_colmajor(k):
    return [(r, c) for c in range(k) for r in range(c + 1)]


def chol_transf(x):
    """D <-> its unconstrained parameters: the upper Cholesky factor U (U'U = D),
    column-major over the upper triangle, diagonal on the log scale
    (what the reference's chol_transf() computes, R/Functions.R:140-158).
    Matrix in -> vector out; vector in -> (D, U)."""
    x = np.asarray(x, float)
    if x.ndim == 2:
        k = x.shape[0]
        U = np.linalg.cholesky(x).T.copy()
        U[np.diag_indices(k)] = np.log(np.diag(U))
        return np.array([U[r, c] for r, c in _upper_tri_colmajor(k)])
    k = int(round((-1 + np.sqrt(1 + 8 * len(x))) / 2))
    U = np.zeros((k, k))
    for v, (r, c) in zip(x, _upper_tri_colmajor(k)):
        U[r, c] = v
    U[np.diag_indices(k)] = np.exp(np.diag(U))
    return U.T @ U, U


def split_thetas(thetas, eng, diag_D):
    """c(betas, D-params, phis, gammas), R/mixed_fit.R:240-247."""
    thetas = np.asarray(thetas, float)
    nD = eng.q if diag_D else eng.npack
    o = 0
    betas = thetas[o:o + eng.p]; o += eng.p
    Dp = thetas[o:o + nD]; o += nD
    phis = thetas[o:o + eng.n_phis] if eng.n_phis else None; o += eng.n_phis
    gammas = thetas[o:o + eng.p_zi] if eng.p_zi else None; o += eng.p_zi
    if o != len(thetas):
        raise ValueError("thetas has the wrong length")
    if diag_D:
        return betas, np.diag(np.exp(Dp)), None, phis, gammas
    D, U = chol_transf(Dp)
    return betas, D, U, phis, gammas


def score_D(S, sum_w, n, D, U, diag_D):
    """d(-loglik)/d(D-params) from S = sum_i w_i E[b b' | y_i] (R/Fit_Funs.R:237-286).

    With M = (c D^-1 - D^-1 S D^-1) / 2 the derivative with respect to a symmetric
    perturbation dD is <M, dD>; the reference assembles the same number through
    deriv_D() and jacobian2().  c is n for the diagonal parameterisation
    (Fit_Funs.R:246) and sum(weights) for the full one (Fit_Funs.R:283)."""
    q = D.shape[0]
    Dinv = np.linalg.inv(D)
    if diag_D:
        dd = np.diag(D)
        return 0.5 * (n - np.diag(S) / dd)
    M = 0.5 * (sum_w * Dinv - Dinv @ S @ Dinv)
    out = []
    for r, c in _upper_tri_colmajor(q):
        dU = np.zeros((q, q))
        dU[r, c] = U[r, c] if r == c else 1.0      # diagonal is exp(theta)
        dD = dU.T @ U + U.T @ dU
        out.append(np.sum(M * dD))
    return np.array(out)


# ----------------------------------------------------------------------------
# GHfun / logLik_mixed / score_mixed
# ----------------------------------------------------------------------------
@dataclass
class GH:
    """What GHfun() returns that the rest of mixed_fit() reads (R/Functions.R:135-137).
    The N x K / nK x q matrices b, b2, Ztb, Z_zitb are never built: the grid lives on
    the device as (post_modes, chol)."""
    engine: Engine
    post_modes: np.ndarray       # n x q
    chol: np.ndarray             # n x q(q+1)/2, chol(post_hessians)[upper.tri(., TRUE)]
    log_dets: np.ndarray         # n
    wGH: np.ndarray              # K
    stats: dict
    version: object = None

    @property
    def post_vars(self):
        """lapply(post_hessians, solve): (R'R)^-1 per cluster."""
        q = self.engine.q
        out = np.empty((len(self.post_modes), q, q))
        for i in range(len(self.post_modes)):
            R = np.zeros((q, q))
            for v, (r, c) in zip(self.chol[i], _upper_tri_colmajor(q)):
                R[r, c] = v
            Ri = np.linalg.inv(R)
            out[i] = Ri @ Ri.T
        return out


def gauss_hermite_wGH(k, q):
    """wGH of R/Functions.R:124-125: 2^(q/2) prod(w) exp(sum z^2), first dimension fastest."""
    x, w = np.polynomial.hermite.hermgauss(k)
    x, w = x[::-1], w[::-1]
    f = w * np.exp(x * x)
    out = np.ones(1)
    for _ in range(q):
        out = (out[:, None] * f[None, :]).reshape(-1)
    return 2 ** (q / 2) * out


def GHfun(engine, b, betas, inv_D, phis=None, gammas=None):
    """R/Functions.R:105-138.  ``b`` is the warm start (n x q) or None to reuse the modes
    the engine already holds (mixed_fit carries post_modes between calls,
    R/mixed_fit.R:100,117,287); zeros give the reference's cold start."""
    warm = True
    if b is not None:
        b = np.asarray(b, float)
        if not np.any(b != 0.0):
            warm = False
        else:
            n_loc = getattr(engine, "n_local", engine.n)      # a ShardedEngine holds its own clusters only
            b = b.reshape(n_loc, engine.q)
            eye = np.zeros((n_loc, engine.npack))
            for t, (r, c) in enumerate(_upper_tri_colmajor(engine.q)):
                eye[:, t] = 1.0 if r == c else 0.0
            engine.set_modes(b, eye)
    stats = engine.find_modes(betas, inv_D, phis, gammas, warm=warm)
    modes, chol, logdet = engine.get_modes()
    return GH(engine, modes, chol, logdet, gauss_hermite_wGH(engine.nAGQ, engine.q), stats,
              version=engine._modes_version)


def engine_X(engine, zi=False):
    """host copy of X / X_zi kept for the per-observation score contributions"""
    return engine._X_zi_host if zi else engine._X_host


def _ensure_grid(engine, gh):
    if gh is not None and gh.version is not engine._modes_version:
        engine.set_modes(gh.post_modes, gh.chol)
        gh.version = engine._modes_version


def _fused_eval(thetas, engine, gh, diag_D):
    _ensure_grid(engine, gh)
    key = (np.asarray(thetas, float).tobytes(), bool(diag_D), id(engine._modes_version))
    if engine._cache_key == key:
        return engine._cache_val
    betas, D, U, phis, gammas = split_thetas(thetas, engine, diag_D)
    val = (engine.eval(betas, D, phis, gammas, want_score=True), D, U)
    engine._cache_key, engine._cache_val = key, val
    return val


class Penalty:
    """Student's t penalty on the fixed effects, `penalized = TRUE` or a list(pen_mu, pen_sigma,
    pen_df) in mixed_model() (R/mixed_model.R:196-208; R/mixed_fit.R:64-69): the log-likelihood gains
    dmvt(betas, pen_mu, invSigma = diag(1 / pen_sigma^2), pen_df, log = TRUE, prop = TRUE)
    (R/Fit_Funs.R:39-40, R/Functions.R:424-462) and the beta-score the term of R/Fit_Funs.R:169-173
    (written there with betas, not betas - pen_mu).  p-sized host algebra: the engine is not involved."""

    def __init__(self, ncx, pen_mu=0.0, pen_sigma=1.0, pen_df=3.0):
        self.mu = np.resize(np.atleast_1d(np.asarray(pen_mu, float)), ncx)
        self.inv_sigma = np.resize(1.0 / np.atleast_1d(np.asarray(pen_sigma, float)) ** 2, ncx)
        self.df = float(pen_df)

    @classmethod
    def from_arg(cls, penalized, ncx):
        if penalized is None or penalized is False:
            return None
        if penalized is True:
            return cls(ncx)
        if isinstance(penalized, cls):
            return penalized
        if isinstance(penalized, dict):
            if not set(penalized) <= {"pen_mu", "pen_sigma", "pen_df"}:
                raise ValueError("when 'penalized' is a dict it needs the components 'pen_mu', 'pen_sigma' and 'pen_df'")
            return cls(ncx, **penalized)
        raise ValueError("argument 'penalized' must be a logical or a dict")

    def log_dens(self, betas):
        d = np.asarray(betas, float) - self.mu
        return -0.5 * (self.df + len(d)) * np.log(1.0 + np.sum(d * self.inv_sigma * d) / self.df)

    def score(self, betas):
        betas = np.asarray(betas, float)
        v = betas * self.inv_sigma / self.df
        return v * ((self.df + len(betas)) / (1.0 + betas @ v))

    def hessian(self, betas):
        """Jacobian of score() (symmetric): the penalty's block of the observed information."""
        betas = np.asarray(betas, float)
        v = betas * self.inv_sigma / self.df
        den = 1.0 + betas @ v
        f = (self.df + len(betas)) / den
        return f * np.diag(self.inv_sigma / self.df) - (2.0 * f / den) * np.outer(v, v)


def logLik_mixed(thetas, engine, gh, diag_D=False, i_contributions=False, fused=True, penalty=None):
    """R/Fit_Funs.R:1-42: the NEGATIVE marginal log-likelihood on the fixed grid ``gh`` (minus the
    log penalty when ``penalty`` is given); with ``i_contributions`` the per-cluster vector
    -w_i log p(y_i).  ``fused`` evaluates the score in the same pass and keeps it for a following
    score_mixed() at the same thetas (optim calls fn then gr)."""
    pen = 0.0 if penalty is None else penalty.log_dens(np.asarray(thetas, float)[:engine.p])
    if i_contributions:
        _ensure_grid(engine, gh)
        betas, D, U, phis, gammas = split_thetas(thetas, engine, diag_D)
        return -engine.eval_contrib(betas, D, phis, gammas)["ll_i"] - pen
    if fused:
        res, _, _ = _fused_eval(thetas, engine, gh, diag_D)
        return -res["loglik"] - pen
    _ensure_grid(engine, gh)
    betas, D, U, phis, gammas = split_thetas(thetas, engine, diag_D)
    return -engine.eval(betas, D, phis, gammas, want_score=False)["loglik"] - pen


def score_mixed(thetas, engine, gh, diag_D=False, i_contributions=False, penalty=None):
    """R/Fit_Funs.R:44-293: gradient of the NEGATIVE log-likelihood (with the penalty term of :169-173),
    c(score.betas, score.D, score.phis, score.gammas).  With ``i_contributions`` the
    per-observation / per-cluster pieces (R/Fit_Funs.R:288-290)."""
    if i_contributions:
        _ensure_grid(engine, gh)
        betas, D, U, phis, gammas = split_thetas(thetas, engine, diag_D)
        c = engine.eval_contrib(betas, D, phis, gammas)
        out = dict(score_betas=-engine_X(engine) * c["zbar"][:, None])
        if diag_D:
            dd = np.diag(D)
            out["score_D"] = 0.5 * (1.0 - np.einsum("idd->id", c["S_i"]) / dd[None, :])
        else:
            out["score_D"] = np.array([score_D(Si, 1.0, 1, D, U, False) for Si in c["S_i"]])
        if engine.n_phis:
            out["score_phis"] = -c["sphi_obs"]
        if engine.p_zi:
            out["score_gammas"] = -engine_X(engine, zi=True) * c["zbar_zi"][:, None]
        return out
    res, D, U = _fused_eval(thetas, engine, gh, diag_D)
    g_beta = -res["g_beta"]
    if penalty is not None:
        g_beta = g_beta + penalty.score(np.asarray(thetas, float)[:engine.p])
    parts = [g_beta, score_D(res["S"], res["sum_w"], engine.n, D, U, diag_D)]
    if engine.n_phis:
        parts.append(-res["g_phi"])
    if engine.p_zi:
        parts.append(-res["g_gamma"])
    return np.concatenate(parts)
