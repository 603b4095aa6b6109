// Host build of glmmadaptive_b200/csrc/fastmath.cuh for tests/test_fastmath.py
// (test infrastructure; the product evaluates these functions on the GPU only).
#include "../glmmadaptive_b200/csrc/fastmath.cuh"
#include "../glmmadaptive_b200/csrc/fastmath_tables.h"

extern "C" {
void hm_log_rcp(const double* t, int n, double* L, double* r) {
    double c[GLMMA_FM_NCONST]; fastmath_constants(c);
    MathCtx cx = make_mathctx(GLMMA_LOGTAB_HOST, GLMMA_EXPTAB_HOST, c);
    for (int i = 0; i < n; ++i) fast_log_rcp(t[i], cx, L[i], r[i]);
}
void hm_log_rcp_vec4(const double* t, int n, double* L, double* r) {
    double c[GLMMA_FM_NCONST]; fastmath_constants(c);
    MathCtx cx = make_mathctx(GLMMA_LOGTAB_HOST, GLMMA_EXPTAB_HOST, c);
    for (int i = 0; i + 4 <= n; i += 4) fast_log_rcp_vec<4>(t + i, cx, L + i, r + i);
}
void hm_exp(const double* x, int n, double* out) {
    double c[GLMMA_FM_NCONST]; fastmath_constants(c);
    MathCtx cx = make_mathctx(GLMMA_LOGTAB_HOST, GLMMA_EXPTAB_HOST, c);
    for (int i = 0; i < n; ++i) out[i] = fast_exp(x[i], cx);
}
void hm_exp_vec3(const double* x, int n, double* out) {
    double c[GLMMA_FM_NCONST]; fastmath_constants(c);
    MathCtx cx = make_mathctx(GLMMA_LOGTAB_HOST, GLMMA_EXPTAB_HOST, c);
    for (int i = 0; i + 3 <= n; i += 3) fast_exp_vec<3>(x + i, cx, out + i);
}
void hm_logphi_mills(const double* x, int n, double* lp, double* ml) {
    double c[GLMMA_FM_NCONST]; fastmath_constants(c);
    MathCtx cx = make_mathctx(GLMMA_LOGTAB_HOST, GLMMA_EXPTAB_HOST, c);
    for (int i = 0; i < n; ++i) fast_logphi_mills(x[i], cx, lp[i], ml[i]);
}
void hm_erfcx(const double* u, int n, double* out) {
    double c[GLMMA_FM_NCONST]; fastmath_constants(c);
    MathCtx cx = make_mathctx(GLMMA_LOGTAB_HOST, GLMMA_EXPTAB_HOST, c);
    for (int i = 0; i < n; ++i) out[i] = fast_erfcx(u[i], cx);
}
void hm_lgamma_digamma(const double* x, int n, double* lg, double* ps) {
    double c[GLMMA_FM_NCONST]; fastmath_constants(c);
    MathCtx cx = make_mathctx(GLMMA_LOGTAB_HOST, GLMMA_EXPTAB_HOST, c);
    for (int i = 0; i < n; ++i) fast_lgamma_digamma(x[i], cx, lg[i], ps[i]);
}
void hm_one_minus_expneg(const double* x, int n, double* out) {
    double c[GLMMA_FM_NCONST]; fastmath_constants(c);
    MathCtx cx = make_mathctx(GLMMA_LOGTAB_HOST, GLMMA_EXPTAB_HOST, c);
    for (int i = 0; i < n; ++i) out[i] = fast_one_minus_expneg(x[i], cx);
}
void hm_l1p_sigmoid(const double* ex, int n, double* l1p, double* sig) {
    double c[GLMMA_FM_NCONST]; fastmath_constants(c);
    MathCtx cx = make_mathctx(GLMMA_LOGTAB_HOST, GLMMA_EXPTAB_HOST, c);
    for (int i = 0; i < n; ++i) fast_l1p_sigmoid(ex[i], cx, l1p[i], sig[i]);
}
}
