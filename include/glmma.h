/*
 * glmma.h -- C ABI of the B200-native engine for GLMMadaptive's AGQ hot path.
 *
 * The reference (drizopoulos/GLMMadaptive v0.9-7) is pure R and has no native
 * boundary; the seam this library sits behind is the set of R closures
 * mixed_fit() calls (paths relative to the reference tree):
 *
 *   GHfun() / find_modes()        R/Functions.R:1-138   (call sites R/mixed_fit.R:89,105,267,315)
 *   logLik_mixed()                R/Fit_Funs.R:1-42     (optim fn,  R/mixed_fit.R:271,319-328)
 *   score_mixed()                 R/Fit_Funs.R:44-293   (optim gr,  R/mixed_fit.R:271,329-342)
 *   family log_dens / score_*     R/Fit_Funs.R:429-1435
 *
 * Every entry point below is what an R .Call shim or any
 * other FFI (ctypes, cgo, JNI) binds: plain pointers and sizes, no C++ or torch
 * types, int status codes (0 = ok), no exceptions across the boundary.
 * All matrices are COLUMN-MAJOR doubles, exactly as R stores them, so the shim
 * passes REAL(x) pointers through unchanged.  Inputs are borrowed for the
 * duration of a call only.
 *
 * There is no CPU fallback: every compute entry point fails with
 * GLMMA_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef GLMMA_H
#define GLMMA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLMMA_VERSION 200

#if defined(__GNUC__)
#define GLMMA_API __attribute__((visibility("default")))
#else
#define GLMMA_API
#endif

/* status codes */
#define GLMMA_OK            0
#define GLMMA_ERR_ARG       1   /* bad argument / unsupported configuration   */
#define GLMMA_ERR_CUDA      2   /* CUDA runtime error (sticky on the handle)  */
#define GLMMA_ERR_STATE     3   /* call order (e.g. eval before modes exist)  */
#define GLMMA_ERR_NUMERIC   4   /* non-PD matrix on the host side (D, invD)   */

/* family$family strings of the reference -> ids (R/Fit_Funs.R, see glmma_family_id) */
enum glmma_family {
    GLMMA_BINOMIAL = 0,          /* stats::binomial()            R/Fit_Funs.R:429-438   */
    GLMMA_POISSON = 1,           /* stats::poisson()             R/Fit_Funs.R:440-445   */
    GLMMA_NEGBIN = 2,            /* negative.binomial()          R/Fit_Funs.R:467-513   */
    GLMMA_ZI_POISSON = 3,        /* zi.poisson()                 R/Fit_Funs.R:515-563   */
    GLMMA_ZI_NEGBIN = 4,         /* zi.negative.binomial()       R/Fit_Funs.R:565-640   */
    GLMMA_ZI_BINOMIAL = 5,       /* zi.binomial()                R/Fit_Funs.R:766-827   */
    GLMMA_HURDLE_POISSON = 6,    /* hurdle.poisson()             R/Fit_Funs.R:642-688   */
    GLMMA_HURDLE_NEGBIN = 7,     /* hurdle.negative.binomial()   R/Fit_Funs.R:690-764   */
    GLMMA_HURDLE_LOGNORMAL = 8,  /* hurdle.lognormal()           R/Fit_Funs.R:829-885   */
    GLMMA_BETA = 9,              /* beta.fam()                   R/Fit_Funs.R:887-930   */
    GLMMA_HURDLE_BETA = 10,      /* hurdle.beta.fam()            R/Fit_Funs.R:932-1004  */
    GLMMA_GAMMA = 11,            /* Gamma.fam()                  R/Fit_Funs.R:1323-1358 */
    GLMMA_CENSORED_NORMAL = 12,  /* censored.normal()            R/Fit_Funs.R:1360-1435 */
    GLMMA_STUDENTS_T = 13,       /* students.t(df)               R/Fit_Funs.R:1006-1041 */
    GLMMA_BETA_BINOMIAL = 14,    /* beta.binomial() (logit, cloglog) R/Fit_Funs.R:1245-1321 */
    GLMMA_N_FAMILIES = 15
};

/* family$link; only the link each family object is built with is accelerated
 * (binomial: logit, probit, cloglog).  Anything else stays on the R path. */
enum glmma_link {
    GLMMA_LINK_DEFAULT = 0,
    GLMMA_LINK_LOGIT = 1,
    GLMMA_LINK_LOG = 2,
    GLMMA_LINK_IDENTITY = 3,
    GLMMA_LINK_PROBIT = 4,
    GLMMA_LINK_CLOGLOG = 5
};

/* What mixed_fit() receives (R/mixed_fit.R:1-25), before it splits it into
 * per-cluster lists.  Rows must be ordered so that each cluster is contiguous
 * (mixed_model() sorts by the grouping factor, R/mixed_model.R:50,104). */
typedef struct glmma_data {
    int32_t family;         /* enum glmma_family                                         */
    int32_t link;           /* enum glmma_link                                           */
    int64_t n_obs;          /* N  = nrow(X)                                              */
    int64_t n_clusters;     /* n  = length(unique(id))                                   */
    int32_t p;              /* ncol(X)                                                   */
    int32_t q_y;            /* ncol(Z);  q = q_y + q_zi <= 4 is compiled in (GLMMA_ERR_ARG beyond) */
    int32_t p_zi;           /* ncol(X_zi), 0 if zi_fixed absent                          */
    int32_t q_zi;           /* ncol(Z_zi), 0 if zi_random absent                         */
    int32_t n_phis;         /* length(phis)                                              */
    int32_t y_cols;         /* NCOL(y): 1, or 2 (binomial successes/failures;
                               censored.normal value/indicator)                          */
    int32_t nAGQ;           /* control$nAGQ                                              */
    int32_t reserved;
    const double* y;        /* N x y_cols                                                */
    const double* X;        /* N x p                                                     */
    const double* Z;        /* N x q_y                                                   */
    const double* X_zi;     /* N x p_zi or NULL                                          */
    const double* Z_zi;     /* N x q_zi or NULL                                          */
    const double* offset;   /* N or NULL                                                 */
    const double* offset_zi;/* N or NULL                                                 */
    const double* weights;  /* n (one per cluster, as mixed_model(weights=)) or NULL     */
    const int32_t* id;      /* N cluster labels, equal values contiguous (R's `id`);
                               used when row_ptr is NULL                                 */
    const int64_t* row_ptr; /* n+1 CSR offsets into the rows, or NULL                    */
    double family_df;       /* students.t: family$df (fixed by the family object); else 0 */
} glmma_data;

/* Per-call statistics of the mode search (replaces the silent behaviour of
 * optim() inside find_modes, R/Functions.R:80-89). */
typedef struct glmma_mode_stats {
    int64_t n_not_converged;   /* clusters that hit the iteration cap                  */
    int64_t n_chol_failed;     /* clusters whose Hessian at the mode is not PD
                                  (the reference's chol() would stop(), Functions.R:115) */
    int32_t max_iterations;    /* largest Newton iteration count over clusters          */
    int32_t reserved;
    double  mean_iterations;
} glmma_mode_stats;

typedef struct glmma_handle glmma_handle;

/* ---- lifetime ---------------------------------------------------------- */

/* Copies the data to `device`, builds the CSR cluster index and the
 * Gauss-Hermite table (gauher(nAGQ), R/Functions.R:306-341).  Modes start at
 * 0 (post_modes <- matrix(0, n, nRE), R/mixed_fit.R:63). */
GLMMA_API int glmma_create(const glmma_data* data, int device, glmma_handle** out);
/* One process, several GPUs (R is single-process): the clusters are split into `ndev` contiguous
 * ranges balanced on observations, one per device in `devices`; every entry point below then works on
 * the returned handle as on a single-device one (results are the sums over the devices, added on the
 * host in device order; per-cluster / per-observation outputs are concatenated).  glmma_eval_device,
 * glmma_set_stream and glmma_time_eval need a single-device handle.  ndev = 1 is glmma_create. */
GLMMA_API int glmma_create_multi(const glmma_data* data, const int* devices, int ndev, glmma_handle** out);
GLMMA_API int glmma_n_devices(const glmma_handle* h);
GLMMA_API void glmma_destroy(glmma_handle* h);
GLMMA_API const char* glmma_last_error(const glmma_handle* h);   /* h may be NULL: last create() error */
GLMMA_API int glmma_version(void);
/* family$family (e.g. "zero-inflated poisson") -> enum glmma_family, -1 if not accelerated */
GLMMA_API int glmma_family_id(const char* family_name);
/* All kernels of this handle are launched on `cuda_stream` (a cudaStream_t;
 * NULL = the legacy default stream). */
GLMMA_API int glmma_set_stream(glmma_handle* h, void* cuda_stream);

/* ---- sizes ------------------------------------------------------------- */
GLMMA_API int64_t glmma_n_clusters(const glmma_handle* h);
GLMMA_API int64_t glmma_n_obs(const glmma_handle* h);
GLMMA_API int     glmma_nre(const glmma_handle* h);          /* q = q_y + q_zi              */
GLMMA_API int     glmma_n_nodes(const glmma_handle* h);      /* K = nAGQ^q                  */
/* length of the result vector of glmma_eval: 4 + p + n_phis + p_zi + q*q */
GLMMA_API int     glmma_result_len(const glmma_handle* h);

/* ---- GHfun / find_modes (R/Functions.R:1-138) -------------------------- */

/* Per-cluster mode of  -sum_j log f(y_ij | b) + b' invD b / 2  (Functions.R:5-20)
 * by damped Newton from the current modes (warm = 1, as mixed_fit carries
 * post_modes between GHfun calls, R/mixed_fit.R:100,117,287) or from 0
 * (warm = 0), then the reference's cd_vec Hessian of the analytic gradient at
 * the mode (Functions.R:90,249-262), its upper Cholesky factor, R^-1 and
 * log_dets = -sum log diag R (Functions.R:115-122).  invD is q x q. */
GLMMA_API int glmma_find_modes(glmma_handle* h, const double* betas, const double* invD,
                     const double* phis, const double* gammas, int warm,
                     glmma_mode_stats* stats /* may be NULL */);

/* GH$post_modes (n x q), chol(post_hessians) as packed upper triangles
 * (n x q(q+1)/2, entry order (0,0),(0,1),(1,1),(0,2),... = R's
 * U[upper.tri(U, TRUE)] per cluster, cluster-major) and GH$log_dets (n).
 * Any output pointer may be NULL. */
GLMMA_API int glmma_get_modes(glmma_handle* h, double* post_modes, double* chol_upper, double* log_dets);
/* Injects modes and Cholesky factors computed elsewhere (tests use the oracle's
 * so that eval parity is independent of the mode search). Same layouts. */
GLMMA_API int glmma_set_modes(glmma_handle* h, const double* post_modes, const double* chol_upper);

/* ---- logLik_mixed + score_mixed (R/Fit_Funs.R:1-293) -------------------- */

/* One fused evaluation on the current grid.  D is the q x q covariance matrix
 * (already back-transformed from thetas by the caller: diag(exp(.)) or
 * chol_transf(.), Fit_Funs.R:10).  Result layout (doubles):
 *   [0] loglik      =  sum_i w_i log p(y_i)        (logLik_mixed returns its negative)
 *   [1] sum_w       =  sum of w_i over the clusters that entered the sums
 *   [2] n_nonfinite =  clusters whose log p(y_i) is NaN/Inf (dropped, as na.rm = TRUE does)
 *   [3] reserved
 *   [4 .. 4+p)                 g_beta  = sum_i w_i sum_k pi_ik sum_j x_ij  s_eta(ijk)
 *   [.. +n_phis)               g_phi   = sum_i w_i sum_k pi_ik sum_j s_phi(ijk)
 *   [.. +p_zi)                 g_gamma = sum_i w_i sum_k pi_ik sum_j xzi_ij s_etazi(ijk)
 *   [.. +q*q)                  S       = sum_i w_i E[b b' | y_i]   (column-major)
 * score_mixed's vector is  c(-g_beta, score.D(S, sum_w, D), -g_phi, -g_gamma)
 * (SURVEY.md section 8a row a12).  With want_score = 0 only [0..3] are computed. */
GLMMA_API int glmma_eval(glmma_handle* h, const double* betas, const double* D, const double* phis,
               const double* gammas, int want_score, double* result);

/* Same, but the result vector is left in DEVICE memory at d_result (on the
 * handle's stream, no synchronisation) so the caller can all-reduce it across
 * ranks (NCCL) before reading it back. */
GLMMA_API int glmma_eval_device(glmma_handle* h, const double* betas, const double* D, const double* phis,
                      const double* gammas, int want_score, double* d_result);

/* i_contributions = TRUE pieces (Fit_Funs.R:34-35, 102-103, 186-187, 228-229, 255-269):
 *   ll_i     (n)      w_i log p(y_i)
 *   zbar     (N)      w_i sum_k pi_ik s_eta(ijk)       -> score.betas[j, l]  = -X[j, l]   * zbar[j]
 *   zbar_zi  (N)      w_i sum_k pi_ik s_etazi(ijk)     -> score.gammas[j, l] = -Xzi[j, l] * zbar_zi[j]
 *   sphi_obs (N)      w_i sum_k pi_ik s_phi(ijk)       -> score.phis[j]      = -sphi_obs[j]
 *   S_i      (n*q*q)  w_i E[b b' | y_i], cluster-major, each q x q column-major
 * Any pointer may be NULL. Host pointers. */
GLMMA_API int glmma_eval_contrib(glmma_handle* h, const double* betas, const double* D, const double* phis,
                       const double* gammas, double* ll_i, double* zbar, double* zbar_zi,
                       double* sphi_obs, double* S_i);

/* ---- EM phase of mixed_fit (R/mixed_fit.R:119-229; score_betas / score_phis /
 *      score_gammas, R/Fit_Funs.R:295-427) ---------------------------------- */

/* E-step on the current grid: the same result vector as glmma_eval(want_score = 1)
 * (loglik = lgLik[it] of R/mixed_fit.R:155; S / n = the D update of :186-190), and the
 * posterior node weights  p_by[i, k] * wGH[k]  (R/mixed_fit.R:141-143) are kept on the
 * device (n x K doubles) for the M-step calls that follow. */
GLMMA_API int glmma_em_estep(glmma_handle* h, const double* betas, const double* D, const double* phis,
                   const double* gammas, double* result);

/* M-step scores with p_by held fixed at the last E-step's value:
 *   g_beta  = -score_betas(betas, ..., p_by)    (R/Fit_Funs.R:295-365)
 *   g_phi   = -score_phis(phis, ..., p_by)      (R/Fit_Funs.R:367-396)
 *   g_gamma = -score_gammas(gammas, ..., p_by)  (R/Fit_Funs.R:398-427)
 * evaluated at the (betas, phis, gammas) passed here, in glmma_eval's result layout
 * ([0] = 0, S = sum_i w_i E_p_by[b b'] on the frozen weights).  mixed_fit() differentiates these numerically
 * for its Newton updates, so one EM iteration makes 3 + 2 * (p + n_phis + p_zi) such calls at most. */
GLMMA_API int glmma_em_score(glmma_handle* h, const double* betas, const double* phis, const double* gammas,
                   double* result);

/* ---- starting values of mixed_model() (R/mixed_model.R:157-194) ------------ */

/* One iteration of stats::glm.fit (IRLS) on the resident data, as one pass over the rows:
 *   design     0: X (ncol p), 1: X_zi (ncol p_zi); at most 8 columns
 *   glm_family 0 binomial(logit), 1 poisson(log), 2 Gamma(inverse), 3 gaussian(identity)
 *   response   0: the model response (two-column binomial: successes / trials with prior weights
 *                 trials), 1: the zero indicator as.numeric(y == 0) (the gammas start, :181-186)
 *   use_offset 1: the design's offset enters (mixed_model passes `offset` / `offset_zi`)
 *   first      1: eta from family$initialize's mustart (beta ignored), 0: eta = X beta + offset
 * out (1 + p + p*p doubles): [0] deviance at the current mu, [1..p] X'Wz, then X'WX (column-major),
 * with W = prior * mu.eta^2 / V(mu), z = (eta - offset) + (y - mu) / mu.eta.  The caller solves the
 * p x p system and applies glm.fit's stopping rule |dev - dev_old| / (|dev| + 0.1) < 1e-8, maxit 25
 * (glmmadaptive_b200/fit.py: glm_fit_device); sums over cluster shards add, so the ranks of a
 * sharded fit all-reduce `out` before solving. */
GLMMA_API int glmma_glm_pass(glmma_handle* h, int design, int glm_family, int response, int use_offset,
                   int first, const double* beta, double* out);

/* Observed information: the Hessian of the NEGATIVE marginal log-likelihood on the current grid with
 * respect to theta = c(betas, D parameters, phis, gammas) (reference order, R/mixed_fit.R:240-247;
 * D parameters: log(diag(D)) when diag_D, else chol_transf(D)), nparams x nparams column-major.  The
 * reference differentiates score_mixed numerically for it (cd_vec, R/mixed_fit.R:324-342: 2 * nparams
 * evaluations, relative truncation error ~1e-6); this is one pass over the data with the exact
 * structure E[d2a] + Cov(da) per cluster and the families' analytic scores differenced point-wise
 * (kernels.cuh: info_kernel), agreeing with cd_vec(score_mixed) to its truncation error.  Supports up to
 * 12 parameters (GLMMA_ERR_ARG beyond: fall back to finite differences of glmma_eval).  The penalty
 * term of penalized = TRUE (R/Fit_Funs.R:169-173) is the caller's (p x p host algebra). */
GLMMA_API int glmma_observed_information(glmma_handle* h, const double* betas, const double* D,
                               const double* phis, const double* gammas, int diag_D, double* H);

/* log_post_b of predict(type = "subject_specific", se.fit = TRUE)'s Metropolis step
 * (reference R/methods.R:1020-1036, called there through mapply once per cluster for the current and
 * once for the proposed random effects of every Monte-Carlo draw, :1093-1112): for ONE vector b_i per
 * cluster, b an n x q column-major matrix,
 *   out[i] = sum_j log f(y_ij | X_ij betas + Z_ij b_i[1:q_y], phis, X_zi,ij gammas + Z_zi,ij b_i[-(1:q_y)])
 *            - b_i' invD b_i / 2            (non-finite log-density terms dropped: na.rm = TRUE)
 * at the given (betas, invD, phis, gammas) -- the parameter draw of that Monte-Carlo iteration. */
GLMMA_API int glmma_log_post_b(glmma_handle* h, const double* b, const double* betas, const double* invD,
                     const double* phis, const double* gammas, double* out);

/* ---- cluster-sharded jobs: one process per GPU, all-reduce inside the engine ----------------
 * Clusters are conditionally independent, so the log-likelihood and every score component are sums over
 * shards (reference R/Fit_Funs.R:33-37,105,194,231,244-247,271-278).  When each GPU of the box is driven
 * by its own process (one handle per process on its shard of the clusters), connecting the handles makes
 * glmma_eval, glmma_eval_device, glmma_em_estep and glmma_em_score return the sum over ALL ranks: the
 * block that finalises a rank's evaluation stores its ~100-byte vector into every rank's exchange
 * buffer through peer memory (NVLink), waits for the others' and adds them in rank order -- the same
 * bits on every rank, no collective launch, no host round trip.  All ranks must then issue the same
 * sequence of these calls (they do: every rank runs the same optimiser on the same reduced numbers).
 *   1. every rank:  glmma_comm_export(h, bytes)           bytes: GLMMA_COMM_HANDLE_BYTES
 *   2. the host program gathers the nranks handles (any transport: MPI, torch.distributed, files)
 *   3. every rank:  glmma_comm_connect(h, rank, nranks, all_handles)     handles in rank order
 * Mode searches, per-cluster contributions and glmma_glm_pass stay local to the rank. */
#define GLMMA_COMM_HANDLE_BYTES 64
GLMMA_API int glmma_comm_export(glmma_handle* h, void* handle_bytes);
GLMMA_API int glmma_comm_connect(glmma_handle* h, int rank, int nranks, const void* all_handles);
GLMMA_API int glmma_comm_size(const glmma_handle* h);      /* ranks summed by an evaluation (1: not connected) */

/* ---- measurement helpers (bench.py; not part of the reference seam) ----- */

/* Runs `iters` evaluations back to back on the handle's stream and returns the
 * average device time of one evaluation and of its dominant kernel (CUDA events). */
GLMMA_API int glmma_time_eval(glmma_handle* h, const double* betas, const double* D, const double* phis,
                    const double* gammas, int want_score, int iters, float* ms_per_eval,
                    float* ms_main_kernel);
/* Dependent-free DFMA microbenchmark on `device`: measured FP64 FMA peak in TFLOP/s. */
GLMMA_API int glmma_measure_fp64_peak(int device, double* tflops);
/* number of kernel launches issued by this handle so far */
GLMMA_API int64_t glmma_launch_count(const glmma_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* GLMMA_H */
