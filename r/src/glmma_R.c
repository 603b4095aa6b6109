/* .Call glue between GLMMadaptive's R closures and the C ABI of the B200 engine (include/glmma.h).
 *
 * Pure marshalling: every function unpacks SEXP arguments into plain pointers, calls ONE glmma_*
 * entry point and wraps the answer in freshly allocated R vectors.  The reference closures each
 * function stands behind (drizopoulos/GLMMadaptive v0.9-7):
 *   glmma_R_create       mixed_fit()'s data intake             R/mixed_fit.R:1-25
 *   glmma_R_find_modes   GHfun() -> find_modes()               R/Functions.R:1-138
 *   glmma_R_set_modes    GH re-use between optim() calls       R/mixed_fit.R:100,117,287
 *   glmma_R_eval         logLik_mixed() / score_mixed()        R/Fit_Funs.R:1-293
 *   glmma_R_eval_contrib i_contributions = TRUE branches       R/Fit_Funs.R:34-35,102-103,186-187,228-229,255-269
 *   glmma_R_em_estep     E-step of the EM phase                R/mixed_fit.R:120-159
 *   glmma_R_em_score     score_betas/phis/gammas, p_by fixed   R/Fit_Funs.R:295-427
 *   glmma_R_glm_pass     one IRLS pass of the glm.fit starts   R/mixed_model.R:162-194
 *   glmma_R_log_post_b   log_post_b of predict()'s MH step     R/methods.R:1020-1036
 * The R side is r/R/engine_b200.R.  Errors: the ABI returns status codes (no C++ exception or
 * longjmp crosses it), so Rf_error() is raised here with no engine frame live.
 *
 * R is not installed in the build image: tests/test_r_glue.py compiles this file against the minimal
 * declarations in r/stub/ (gcc -fsyntax-only -Wall -Werror, then to an object whose undefined symbols
 * must be exactly R's API plus symbols libglmma.so exports), which pins the glue to include/glmma.h. */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "glmma.h"

static glmma_handle* handle_of(SEXP p) {
    glmma_handle* h = (glmma_handle*) R_ExternalPtrAddr(p);
    if (!h) Rf_error("glmma: the engine handle has been released");
    return h;
}
static void finalizer(SEXP p) {
    glmma_handle* h = (glmma_handle*) R_ExternalPtrAddr(p);
    if (h) glmma_destroy(h);
    R_ClearExternalPtr(p);
}
static const double* opt_real(SEXP x) { return Rf_isNull(x) ? NULL : REAL(x); }

/* link ids of include/glmma.h from family$link */
static int link_id(SEXP link) {
    const char* s = CHAR(STRING_ELT(link, 0));
    static const struct { const char* name; int id; } tab[] = {
        {"logit", GLMMA_LINK_LOGIT}, {"log", GLMMA_LINK_LOG}, {"identity", GLMMA_LINK_IDENTITY},
        {"probit", GLMMA_LINK_PROBIT}, {"cloglog", GLMMA_LINK_CLOGLOG}};
    for (unsigned i = 0; i < sizeof(tab) / sizeof(tab[0]); ++i) {
        const char *a = s, *b = tab[i].name;
        while (*a && *a == *b) { ++a; ++b; }
        if (*a == 0 && *b == 0) return tab[i].id;
    }
    return -1;
}

/* returns NULL (R_NilValue) when the family / link is not accelerated: the caller stays on the R path */
SEXP glmma_R_create(SEXP y, SEXP X, SEXP Z, SEXP X_zi, SEXP Z_zi, SEXP id, SEXP offset, SEXP offset_zi,
                    SEXP weights, SEXP family, SEXP link, SEXP n_phis, SEXP nAGQ, SEXP df, SEXP devices) {
    glmma_data d;
    memset(&d, 0, sizeof(d));
    d.family = glmma_family_id(CHAR(STRING_ELT(family, 0)));
    d.link = link_id(link);
    if (d.family < 0 || d.link < 0) return R_NilValue;
    d.n_obs = Rf_nrows(X);
    d.n_clusters = 0;                       /* derived from id */
    d.p = Rf_ncols(X);
    d.q_y = Rf_ncols(Z);
    d.p_zi = Rf_isNull(X_zi) ? 0 : Rf_ncols(X_zi);
    d.q_zi = Rf_isNull(Z_zi) ? 0 : Rf_ncols(Z_zi);
    d.n_phis = Rf_asInteger(n_phis);
    d.y_cols = Rf_isMatrix(y) ? Rf_ncols(y) : 1;
    d.nAGQ = Rf_asInteger(nAGQ);
    d.family_df = Rf_isNull(df) ? 0.0 : Rf_asReal(df);
    d.y = REAL(y); d.X = REAL(X); d.Z = REAL(Z);
    d.X_zi = opt_real(X_zi); d.Z_zi = opt_real(Z_zi);
    d.offset = opt_real(offset); d.offset_zi = opt_real(offset_zi); d.weights = opt_real(weights);
    d.id = INTEGER(id);
    d.row_ptr = NULL;
    glmma_handle* h = NULL;
    const int ndev = Rf_length(devices);
    const int rc = ndev > 1 ? glmma_create_multi(&d, INTEGER(devices), ndev, &h)
                            : glmma_create(&d, INTEGER(devices)[0], &h);
    if (rc == GLMMA_ERR_ARG && h == NULL) {
        /* dimensions the engine was not compiled for (e.g. q > 3): not an error, the R path handles it */
        Rf_warning("glmma: %s; using the R implementation", glmma_last_error(NULL));
        return R_NilValue;
    }
    if (rc != GLMMA_OK) Rf_error("glmma_create: %s", glmma_last_error(NULL));
    SEXP p = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

SEXP glmma_R_destroy(SEXP p) {
    finalizer(p);
    return R_NilValue;
}

/* list(post_modes n x q, chol q(q+1)/2 x n, log_dets n, c(n_not_converged, n_chol_failed, max_iter, mean_iter)) */
SEXP glmma_R_find_modes(SEXP p, SEXP betas, SEXP invD, SEXP phis, SEXP gammas, SEXP warm) {
    glmma_handle* h = handle_of(p);
    glmma_mode_stats st;
    if (glmma_find_modes(h, REAL(betas), REAL(invD), opt_real(phis), opt_real(gammas), Rf_asLogical(warm), &st))
        Rf_error("glmma_find_modes: %s", glmma_last_error(h));
    const int n = (int) glmma_n_clusters(h), q = glmma_nre(h);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SET_VECTOR_ELT(out, 0, Rf_allocMatrix(REALSXP, n, q));
    SET_VECTOR_ELT(out, 1, Rf_allocMatrix(REALSXP, q * (q + 1) / 2, n));
    SET_VECTOR_ELT(out, 2, Rf_allocVector(REALSXP, n));
    SET_VECTOR_ELT(out, 3, Rf_allocVector(REALSXP, 4));
    if (glmma_get_modes(h, REAL(VECTOR_ELT(out, 0)), REAL(VECTOR_ELT(out, 1)), REAL(VECTOR_ELT(out, 2))))
        Rf_error("glmma_get_modes: %s", glmma_last_error(h));
    double* s = REAL(VECTOR_ELT(out, 3));
    s[0] = (double) st.n_not_converged; s[1] = (double) st.n_chol_failed;
    s[2] = (double) st.max_iterations; s[3] = st.mean_iterations;
    UNPROTECT(1);
    return out;
}

SEXP glmma_R_set_modes(SEXP p, SEXP post_modes, SEXP chol) {
    glmma_handle* h = handle_of(p);
    if (glmma_set_modes(h, REAL(post_modes), REAL(chol))) Rf_error("glmma_set_modes: %s", glmma_last_error(h));
    return R_NilValue;
}

/* the result vector of include/glmma.h: [loglik, sum_w, n_nonfinite, 0, g_beta, g_phi, g_gamma, S] */
SEXP glmma_R_eval(SEXP p, SEXP betas, SEXP D, SEXP phis, SEXP gammas, SEXP want_score) {
    glmma_handle* h = handle_of(p);
    SEXP res = PROTECT(Rf_allocVector(REALSXP, glmma_result_len(h)));
    memset(REAL(res), 0, sizeof(double) * (size_t) glmma_result_len(h));
    if (glmma_eval(h, REAL(betas), REAL(D), opt_real(phis), opt_real(gammas), Rf_asLogical(want_score), REAL(res)))
        Rf_error("glmma_eval: %s", glmma_last_error(h));
    UNPROTECT(1);
    return res;
}

/* list(ll_i n, zbar N, zbar_zi N | NULL, sphi_obs N | NULL, S_i q*q x n) */
SEXP glmma_R_eval_contrib(SEXP p, SEXP betas, SEXP D, SEXP phis, SEXP gammas) {
    glmma_handle* h = handle_of(p);
    const int n = (int) glmma_n_clusters(h), q = glmma_nre(h);
    const R_xlen_t N = (R_xlen_t) glmma_n_obs(h);
    const int has_zi = !Rf_isNull(gammas), has_phi = !Rf_isNull(phis);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
    SET_VECTOR_ELT(out, 0, Rf_allocVector(REALSXP, n));
    SET_VECTOR_ELT(out, 1, Rf_allocVector(REALSXP, N));
    if (has_zi) SET_VECTOR_ELT(out, 2, Rf_allocVector(REALSXP, N));
    if (has_phi) SET_VECTOR_ELT(out, 3, Rf_allocVector(REALSXP, N));
    SET_VECTOR_ELT(out, 4, Rf_allocMatrix(REALSXP, q * q, n));
    if (glmma_eval_contrib(h, REAL(betas), REAL(D), opt_real(phis), opt_real(gammas), REAL(VECTOR_ELT(out, 0)),
                           REAL(VECTOR_ELT(out, 1)), has_zi ? REAL(VECTOR_ELT(out, 2)) : NULL,
                           has_phi ? REAL(VECTOR_ELT(out, 3)) : NULL, REAL(VECTOR_ELT(out, 4))))
        Rf_error("glmma_eval_contrib: %s", glmma_last_error(h));
    UNPROTECT(1);
    return out;
}

SEXP glmma_R_em_estep(SEXP p, SEXP betas, SEXP D, SEXP phis, SEXP gammas) {
    glmma_handle* h = handle_of(p);
    SEXP res = PROTECT(Rf_allocVector(REALSXP, glmma_result_len(h)));
    if (glmma_em_estep(h, REAL(betas), REAL(D), opt_real(phis), opt_real(gammas), REAL(res)))
        Rf_error("glmma_em_estep: %s", glmma_last_error(h));
    UNPROTECT(1);
    return res;
}

SEXP glmma_R_em_score(SEXP p, SEXP betas, SEXP phis, SEXP gammas) {
    glmma_handle* h = handle_of(p);
    SEXP res = PROTECT(Rf_allocVector(REALSXP, glmma_result_len(h)));
    if (glmma_em_score(h, REAL(betas), opt_real(phis), opt_real(gammas), REAL(res)))
        Rf_error("glmma_em_score: %s", glmma_last_error(h));
    UNPROTECT(1);
    return res;
}

/* c(deviance, X'Wz (ncol), X'WX (ncol x ncol)); beta = NULL: the first pass (family$initialize starts) */
SEXP glmma_R_glm_pass(SEXP p, SEXP design, SEXP ncol, SEXP glm_family, SEXP response, SEXP use_offset, SEXP beta) {
    glmma_handle* h = handle_of(p);
    const int nc = Rf_asInteger(ncol);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 1 + nc + nc * nc));
    if (glmma_glm_pass(h, Rf_asInteger(design), Rf_asInteger(glm_family), Rf_asInteger(response),
                       Rf_asLogical(use_offset), Rf_isNull(beta), opt_real(beta), REAL(out)))
        Rf_error("glmma_glm_pass: %s", glmma_last_error(h));
    UNPROTECT(1);
    return out;
}

/* b: n x q matrix (one proposal per cluster) -> numeric(n) */
SEXP glmma_R_log_post_b(SEXP p, SEXP b, SEXP betas, SEXP invD, SEXP phis, SEXP gammas) {
    glmma_handle* h = handle_of(p);
    const int n = (int) glmma_n_clusters(h);
    if (Rf_nrows(b) != n || Rf_ncols(b) != glmma_nre(h)) Rf_error("glmma_log_post_b: b must be n x q");
    SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
    if (glmma_log_post_b(h, REAL(b), REAL(betas), REAL(invD), opt_real(phis), opt_real(gammas), REAL(out)))
        Rf_error("glmma_log_post_b: %s", glmma_last_error(h));
    UNPROTECT(1);
    return out;
}

SEXP glmma_R_info(SEXP p) {
    glmma_handle* h = handle_of(p);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 6));
    double* v = REAL(out);
    v[0] = (double) glmma_n_clusters(h); v[1] = (double) glmma_n_obs(h); v[2] = glmma_nre(h);
    v[3] = glmma_n_nodes(h); v[4] = glmma_n_devices(h); v[5] = (double) glmma_launch_count(h);
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"glmma_R_create", (DL_FUNC) &glmma_R_create, 15},
    {"glmma_R_destroy", (DL_FUNC) &glmma_R_destroy, 1},
    {"glmma_R_find_modes", (DL_FUNC) &glmma_R_find_modes, 6},
    {"glmma_R_set_modes", (DL_FUNC) &glmma_R_set_modes, 3},
    {"glmma_R_eval", (DL_FUNC) &glmma_R_eval, 6},
    {"glmma_R_eval_contrib", (DL_FUNC) &glmma_R_eval_contrib, 5},
    {"glmma_R_em_estep", (DL_FUNC) &glmma_R_em_estep, 5},
    {"glmma_R_em_score", (DL_FUNC) &glmma_R_em_score, 4},
    {"glmma_R_glm_pass", (DL_FUNC) &glmma_R_glm_pass, 7},
    {"glmma_R_log_post_b", (DL_FUNC) &glmma_R_log_post_b, 6},
    {"glmma_R_info", (DL_FUNC) &glmma_R_info, 1},
    {NULL, NULL, 0}};

void R_init_GLMMadaptive(DllInfo* dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
