// Family functors: the device-side form of the reference's family closures
// log_dens / score_eta_fun / score_phis_fun / score_eta_zi_fun
// (reference R/Fit_Funs.R:429-1435; formulas tabulated in SURVEY.md appendix B).
//
// Every functor splits log f(y | eta, phi, eta_zi) into a part that depends on
// the quadrature node (through eta / eta_zi) and a part that does not:
//
//   log f      = point().ld + obs_const().c_ll
//   d/d phis   = point().sp + obs_const().c_phi
//
// The node-independent parts (lgamma / digamma / lchoose terms) cancel in the
// posterior node weights, so they are evaluated once per observation by the
// prep kernel instead of once per node.
//
// y1 / y2 are family-specific transforms of the response prepared once at
// glmma_create() (engine.cu, transform_response()):
//   binomial, zi.binomial : y1 = successes, y2 = trials
//   count families        : y1 = y
//   Gamma, hurdle.lognormal: y1 = y, y2 = log y
//   beta, hurdle.beta     : y1 = log y (-inf when y == 0), y2 = log(1 - y)
//   censored.normal       : y1 = value, y2 = indicator (0 none, 1 left, 2 right)
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

// ----- link pieces with R's clamps (stats::make.link) ------
// The kernels hand every functor emu = exp(eta) and elam = exp(eta_zi) (from the
// per-cluster separable tables, or fast_exp): the functors never call exp(eta).
// log link: mu = pmax(exp(eta), eps); log mu follows the clamp
// CAREFUL = false: the kernel has checked that no clamp is active anywhere on the cluster's grid
// and that every argument of fast_log_rcp / fast_l1p_sigmoid is inside its fast range.
template <bool CAREFUL>
__device__ __forceinline__ void link_log(double eta, double emu, double& mu, double& lmu) {
    mu = emu; lmu = eta;
    if (CAREFUL && eta < GLMMA_LOG_DBL_EPS) { mu = GLMMA_DBL_EPS; lmu = GLMMA_LOG_DBL_EPS; }
}
// logit link: t = exp(eta) clamped to [eps, 1/eps] outside +-30; mu = t / (1 + t).
// Returns mu, log t and log(1 + t); log mu = lt - L, log(1 - mu) = -L.
template <bool CAREFUL>
__device__ __forceinline__ void link_logit(double eta, double emu, const MathCtx& cx, double& mu, double& lt, double& L) {
    double t = emu;
    lt = eta;
    if (CAREFUL && fabs(eta) > 30.0) {
        t = eta < 0.0 ? GLMMA_DBL_EPS : GLMMA_INV_DBL_EPS;
        lt = eta < 0.0 ? GLMMA_LOG_DBL_EPS : -GLMMA_LOG_DBL_EPS;
    }
    double r;
    fast_log_rcp<CAREFUL>(1.0 + t, cx, L, r);
    mu = t * r;
}

// log(1 + lambda) and plogis(eta_zi) from elam = exp(eta_zi)
template <bool CAREFUL>
__device__ __forceinline__ void zero_part(double ez, double elam, const MathCtx& cx, double& l1p, double& pi) {
    fast_l1p_sigmoid<CAREFUL>(elam, cx, l1p, pi);
}

// zl1p / zpi: log(1 + exp(eta_zi)) and plogis(eta_zi) when the caller has them already (ZT = true: the
// register variant tabulates them per (observation, zero-part node index) -- eta_zi depends on the
// node only through the zero-part random effect, so with q_zi <= 1 there are n_i * nAGQ distinct
// values per cluster instead of n_i * nAGQ^q)
// YC (register variant, zero part hoisted -- kernels.cuh, "ZP" mode): the kernel has sorted the cluster's
// observations into positive and zero responses and adds the zero-part terms that every observation
// shares, -log(1 + exp(eta_zi)) in log f and -plogis(eta_zi) in the eta_zi-score, once per NODE from a
// per-cluster table.  point<., false, true, YC> then returns ld and sz WITHOUT those two terms, for an
// observation known to be positive (YC = 1) or zero (YC = 2); zl1p / zpi are not read.  The hurdle
// families have nothing left to evaluate for a zero response (ld = eta_zi, sz = 1: ZERO_FREE), so the
// kernel never calls them with YC = 2.
#define GLMMA_POINT_ARGS double eta, double ez, double emu, double elam, const FamPar& fp, const MathCtx& cx, \
                         double& ld, double& se, double& sz, double& sp, double zl1p = 0.0, double zpi = 0.0

struct FamBase {
    static constexpr bool HAS_ZI = false;    // family has a zero part (eta_zi)
    static constexpr bool HAS_PHI = false;   // family has a dispersion parameter
    static constexpr bool USES_Y2 = false;
    static constexpr bool NEED_EMU = true;   // point() reads emu = exp(eta)
    // range of exp(eta) over a cluster's grid inside which point<., CAREFUL = false> is exact:
    // no link clamp is active and every fast_log_rcp argument is in its fast range
    static constexpr double EMU_LO = 2.3e-16;     // log link: eta >= log(.Machine$double.eps)
    static constexpr double EMU_HI = 1e290;
    // Register-variant (CAREFUL = false) contract:
    //  LIN_ETA: log f contains the term y1 * eta; the kernel adds sum_j y1_j eta_jk once per
    //           node (from sum_j y1_j xb_j and Z'y1) and point<., false> leaves it out.
    //  stage(): per-observation transform of (y1, y2) applied when the record is staged.
    //  finish_se(): point<., false> may return a cheaper carrier c of the eta-score whose
    //           posterior mean maps to the score mean by an affine per-observation map
    //           (e.g. mu instead of y - mu); finish_se(mean(c), y1, y2) undoes it.
    //  sp_scale(): factor applied once per cluster to the accumulated point<., false>().sp
    //  FAST_KIND: how the register variant evaluates a chunk of observations (stage by stage, the
    //           chunk's dependency chains interleaved):
    //           0  generic: point<., false> per observation;
    //           1  t = (prod of the separable factors) + fast_addend(); (L, r) = (log t, 1/t) for the
    //              whole chunk (fast_log_rcp_vec); log f = -rec1 L (+ y eta at node level) with rec1
    //              the staged second record field; the eta-score carrier is r; with HAS_PHI the
    //              phi-score is sp_scale() * (L + rec1 r), whose second term needs no per-point work
    //              (it is sum_j rec1_j * [posterior sum of r_j], formed once per cluster);
    //           2  exp(eta) only: log f = -emu (+ y eta), carrier emu.
    //           3  censored.normal: Gaussian part hoisted to one quadratic form per node, censored
    //              observations from compact records (kernels.cuh).
    //           fast_score(mean(carrier), y1, rec1) maps the posterior mean of the carrier to the
    //           posterior mean of the eta-score.
    static constexpr bool LIN_ETA = false;
    static constexpr bool Y_TABLE = false;       // obs_const() depends on an integer-valued y1 only
    static constexpr int FAST_KIND = 0;
    // log of [EMU_LO, EMU_HI] rounded inwards (range test on the linear predictor)
    static constexpr double LOG_EMU_LO = -36.008;
    static constexpr double LOG_EMU_HI = 667.7;
    // families that do not read exp(eta) can still restrict the register variant to clusters whose
    // linear predictor stays inside [ETA_LO, ETA_HI] on the whole grid (a link clamp outside)
    static constexpr bool HAS_ETA_RANGE = false;
    static constexpr double ETA_LO = -1e300, ETA_HI = 1e300;
    // generic kernel: observations evaluated per trip of its point loop (independent chains for the FP64
    // pipe); 2 for the functors whose careful form is long (beta, censored normal, probit / cloglog)
    static constexpr int GENERIC_UNROLL = 4;
    static constexpr bool ZERO_FREE = false;     // hurdle families: a zero response involves the zero part only
    __device__ static __forceinline__ bool is_pos(double y1) { return y1 > 0.0; }
    __device__ static __forceinline__ double fast_addend(const FamPar&) { return 0.0; }
    __device__ static __forceinline__ double fast_score(double c, double, double, const FamPar&) { return c; }
    __device__ static __forceinline__ void stage(double&, double&, const FamPar&) {}
    __device__ static __forceinline__ double finish_se(double c, double, double, const FamPar&) { return c; }
    __device__ static __forceinline__ double sp_scale(const FamPar&) { return 1.0; }
    __device__ static __forceinline__ void obs_const(double, double, const FamPar&, double& cll, double& cphi) {
        cll = 0.0; cphi = 0.0;
    }
};

// ----- stats::binomial(), logit (canonical) -- R/Fit_Funs.R:429-438, :130-149 --
struct FamBinomialLogit : FamBase {
    static constexpr double EMU_LO = 1e-13;      // logit link: |eta| <= 30 - margin
    static constexpr double EMU_HI = 1e13;
    static constexpr double LOG_EMU_LO = -29.93;
    static constexpr double LOG_EMU_HI = 29.93;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double y, double N, const FamPar&, double& cll, double& cphi) {
        cll = (y == 0.0 || y == N) ? 0.0 : lgamma(N + 1.0) - lgamma(y + 1.0) - lgamma(N - y + 1.0);
        cphi = 0.0;
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double N, GLMMA_POINT_ARGS) {
        double mu, lt, L;
        link_logit<CAREFUL>(eta, emu, cx, mu, lt, L);
        // y log p + (N - y) log q with log p = lt - L, log q = -L
        if (CAREFUL) {
            ld = fma(y, lt, -N * L);
            if (SCORE) se = fma(-N, mu, y);
        } else {
            ld = -N * L;                 // + y eta at node level
            if (SCORE) se = mu;          // carrier: score = y - N mean(mu)
        }
        sz = 0.0; sp = 0.0;
    }
    static constexpr bool LIN_ETA = true;
    __device__ static __forceinline__ double finish_se(double c, double y, double N, const FamPar&) { return fma(-N, c, y); }
    // t = 1 + exp(eta): log f = y eta - N log t, mu = 1 - 1/t
    static constexpr int FAST_KIND = 1;
    __device__ static __forceinline__ double fast_addend(const FamPar&) { return 1.0; }
    __device__ static __forceinline__ double fast_score(double mean_r, double y, double N, const FamPar&) {
        return fma(-N, 1.0 - mean_r, y);
    }
};

// ----- stats::binomial(), probit / cloglog -- score via (y - N mu) mu'(eta) / V(mu),
// R/Fit_Funs.R:150-167 ---------------------------------------------------------
template <int LINK>   // 4 probit, 5 cloglog (enum glmma_link)
struct FamBinomialOther : FamBase {
    static constexpr int GENERIC_UNROLL = 2;
    static constexpr bool USES_Y2 = true;
    // cloglog reads exp(eta) from the separable tables; probit works on eta itself
    static constexpr bool NEED_EMU = LINK == 5;
    // Register variant where no clamp of R's make.link() is active on the grid:
    //   probit : |eta| <= 8 (thresh = -qnorm(eps) = 8.126, dnorm(eta) >= eps)
    //   cloglog: eps <= mu <= 1 - eps and mu.eta >= eps  <=>  1e-15 <= exp(eta) <= 35
    // It evaluates the reference's own expressions -- mu, then q = 1 - mu from the ROUNDED mu, then
    // log mu, log q and (y - N mu) mu.eta / (mu q) -- with the table-based primitives, so that where
    // the reference is ill conditioned (q below 1e-6 depends on the last bits of mu) both variants of
    // the engine are ill conditioned in the same way.
    static constexpr bool HAS_ETA_RANGE = LINK == 4;
    static constexpr double ETA_LO = -8.0, ETA_HI = 8.0;
    static constexpr double EMU_LO = 1e-15, EMU_HI = 35.0;
    static constexpr double LOG_EMU_LO = -34.5, LOG_EMU_HI = 3.55;
    __device__ static __forceinline__ void obs_const(double y, double N, const FamPar& fp, double& cll, double& cphi) {
        FamBinomialLogit::obs_const(y, N, fp, cll, cphi);
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double N, GLMMA_POINT_ARGS) {
        sz = 0.0; sp = 0.0;
        // The generic kernel (CAREFUL) takes the table-based form too wherever no clamp of make.link() is active
        // at this point -- the library erfc / expm1 / log1p cost ~600 instructions per point and made clusters of
        // more than 16 observations 10x slower per point than the register variant
        const bool no_clamp = LINK == 4 ? fabs(eta) <= ETA_HI : (eta >= LOG_EMU_LO && eta <= LOG_EMU_HI);
        if (!CAREFUL || no_clamp) {
            double mu, dmu;
            if (LINK == 4) {
                // Phi(-|eta|) = erfcx(|eta| / sqrt2) exp(-eta^2 / 2) / 2
                const double c = fast_erfcx(fabs(eta) * 0.70710678118654752, cx);
                const double hx = -0.5 * eta, h = hx * eta;
                double E = fast_exp(h, cx);
                E = fma(E, fma(hx, eta, -h), E);
                const double tail = 0.5 * E * c;
                mu = eta >= 0.0 ? 1.0 - tail : tail;
                dmu = 0.39894228040143268 * E;
            } else {
                // mu = 1 - exp(-e^eta), mu.eta = e^eta exp(-e^eta)
                mu = fast_one_minus_expneg(emu, cx);
                dmu = emu * fast_exp(-emu, cx);
            }
            const double q = 1.0 - mu;
            double Lp, rp, Lq, rq;
            fast_log_rcp<false>(mu, cx, Lp, rp);
            fast_log_rcp<false>(q, cx, Lq, rq);
            ld = fma(y, Lp, (N - y) * Lq);
            if (SCORE) se = fma(-N, mu, y) * dmu * (rp * rq);
            else se = 0.0;
            return;
        }
        double mu, dmu;
        if (LINK == 4) {
            const double thresh = 8.125890664701906;   // -qnorm(.Machine$double.eps)
            const double e = fmin(fmax(eta, -thresh), thresh);
            mu = glmma_ndtr(e);
            dmu = fmax(GLMMA_INV_SQRT_2PI * exp(-0.5 * eta * eta), GLMMA_DBL_EPS);
        } else {
            mu = fmax(fmin(-expm1(-exp(eta)), 1.0 - GLMMA_DBL_EPS), GLMMA_DBL_EPS);
            const double e = fmin(eta, 700.0);
            dmu = fmax(exp(e) * exp(-exp(e)), GLMMA_DBL_EPS);
        }
        const double q = 1.0 - mu;
        const double lq = (mu < 0.1) ? log1p(-mu) : log(q);
        const double lp = (q < 0.1) ? log1p(-q) : log(mu);
        ld = (y == 0.0 ? 0.0 : y * lp) + (N - y == 0.0 ? 0.0 : (N - y) * lq);
        if (SCORE) se = (y - N * mu) * dmu / (mu * q);
    }
};

// ----- stats::poisson(), log -- R/Fit_Funs.R:440-445 ---------------------------
struct FamPoisson : FamBase {
    static constexpr bool Y_TABLE = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar&, double& cll, double& cphi) {
        cll = -lgamma(y + 1.0); cphi = 0.0;
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double, GLMMA_POINT_ARGS) {
        double mu, lmu;
        link_log<CAREFUL>(eta, emu, mu, lmu);
        if (CAREFUL) {
            ld = fma(y, lmu, -mu);
            if (SCORE) se = y - mu;
        } else {
            ld = -mu;                    // + y eta at node level
            if (SCORE) se = mu;          // carrier: score = y - mean(mu)
        }
        sz = 0.0; sp = 0.0;
    }
    static constexpr bool LIN_ETA = true;
    __device__ static __forceinline__ double finish_se(double c, double y, double, const FamPar&) { return y - c; }
    static constexpr int FAST_KIND = 2;
    __device__ static __forceinline__ double fast_score(double mean_mu, double y, double, const FamPar&) { return y - mean_mu; }
};

// FamPar layout shared by the negative-binomial type families:
//   a[0] = phi = exp(phis), a[1] = log phi, a[2] = lgamma(phi), a[3] = digamma(phi)
__device__ __forceinline__ void nb_const(double y, const FamPar& fp, double& cll, double& cphi) {
    const double phi = fp.a[0];
    cll = lgamma(y + phi) - fp.a[2] - lgamma(y + 1.0) + phi * fp.a[1];
    cphi = phi * (fp.a[1] + 1.0 - fp.a[3] + glmma_digamma(y + phi));
}

// ----- negative.binomial() -- R/Fit_Funs.R:467-513 -----------------------------
struct FamNegBin : FamBase {
    static constexpr bool Y_TABLE = true;
    static constexpr bool HAS_PHI = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar& fp, double& cll, double& cphi) {
        nb_const(y, fp, cll, cphi);
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double yp_, GLMMA_POINT_ARGS) {
        const double phi = fp.a[0];
        double mu, lmu, L, r;
        link_log<CAREFUL>(eta, emu, mu, lmu);
        fast_log_rcp<CAREFUL>(mu + phi, cx, L, r);
        if (CAREFUL) {
            const double yp = y + phi;
            ld = fma(y, lmu, -yp * L);
            if (SCORE) {
                se = fma(-yp, mu * r, y);
                sp = -phi * fma(yp, r, L);
            }
        } else {
            // staged record: yp_ = y + phi.  log f = y eta (node level) - yp log(mu + phi);
            // with q = yp / (mu + phi): score_eta = phi (q - 1), score_phi = -phi (L + q)
            ld = -yp_ * L;
            if (SCORE) {
                const double q = yp_ * r;
                se = q;                  // carrier, see finish_se
                sp = L + q;              // times sp_scale = -phi at cluster level
            }
        }
        sz = 0.0;
    }
    static constexpr bool LIN_ETA = true;
    __device__ static __forceinline__ void stage(double& y1, double& y2, const FamPar& fp) { y2 = y1 + fp.a[0]; }
    __device__ static __forceinline__ double finish_se(double c, double, double, const FamPar& fp) { return fp.a[0] * (c - 1.0); }
    __device__ static __forceinline__ double sp_scale(const FamPar& fp) { return -fp.a[0]; }
    // t = exp(eta) + phi, rec1 = y + phi: log f = y eta - rec1 log t; score_eta = phi (rec1 / t - 1);
    // score_phi = -phi (log t + rec1 / t) + node-independent part
    static constexpr int FAST_KIND = 1;
    __device__ static __forceinline__ double fast_addend(const FamPar& fp) { return fp.a[0]; }
    __device__ static __forceinline__ double fast_score(double mean_r, double, double yp_, const FamPar& fp) {
        return fp.a[0] * fma(yp_, mean_r, -1.0);
    }
};

// ----- zi.poisson() -- R/Fit_Funs.R:515-563 ------------------------------------
struct FamZiPoisson : FamBase {
    static constexpr bool Y_TABLE = true;
    static constexpr bool LIN_ETA = true;
    static constexpr bool HAS_ZI = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar&, double& cll, double& cphi) {
        cll = (y > 0.0) ? -lgamma(y + 1.0) : 0.0; cphi = 0.0;
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double, GLMMA_POINT_ARGS) {
        double mu, lmu, pi, l1p;
        link_log<CAREFUL>(eta, emu, mu, lmu);
        if (YC != 0) { l1p = 0.0; pi = 0.0; }
        else if (ZT) { l1p = zl1p; pi = zpi; } else zero_part<CAREFUL>(ez, elam, cx, l1p, pi);   // log(1 + lambda), plogis(eta_zi)
        if (YC == 1 || (YC == 0 && y > 0.0)) {
            ld = (CAREFUL ? fma(y, lmu, -mu) : -mu) - l1p;     // fast variant: + y eta at node level
            if (SCORE) { se = y - mu; sz = -pi; }
        } else {
            // log(lambda + exp(-mu)) - log(1 + lambda)
            const double E = fast_exp(-mu, cx);
            double lu, ru;
            fast_log_rcp<CAREFUL>(elam + E, cx, lu, ru);
            ld = lu - l1p;
            if (SCORE) {
                se = -mu * E * ru;
                sz = fma(elam, ru, -pi);
            }
        }
        sp = 0.0;
    }
};

// ----- zi.negative.binomial() -- R/Fit_Funs.R:565-640 --------------------------
struct FamZiNegBin : FamBase {
    static constexpr bool Y_TABLE = true;
    static constexpr bool LIN_ETA = true;
    static constexpr bool HAS_ZI = true;
    static constexpr bool HAS_PHI = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar& fp, double& cll, double& cphi) {
        if (y > 0.0) nb_const(y, fp, cll, cphi); else { cll = 0.0; cphi = 0.0; }
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double, GLMMA_POINT_ARGS) {
        const double phi = fp.a[0];
        double mu, lmu, L, r, pi, l1p;
        link_log<CAREFUL>(eta, emu, mu, lmu);
        fast_log_rcp<CAREFUL>(mu + phi, cx, L, r);
        if (YC != 0) { l1p = 0.0; pi = 0.0; }
        else if (ZT) { l1p = zl1p; pi = zpi; } else zero_part<CAREFUL>(ez, elam, cx, l1p, pi);
        if (YC == 1 || (YC == 0 && y > 0.0)) {
            const double yp = y + phi;
            ld = (CAREFUL ? fma(y, lmu, -yp * L) : -yp * L) - l1p;     // fast variant: + y eta at node level
            if (SCORE) {
                se = fma(-yp, mu * r, y);
                sp = -phi * fma(yp, r, L);
                sz = -pi;
            }
        } else {
            const double lt = fp.a[1] - L;                 // log(phi / (phi + mu))
            const double tphi = fast_exp(phi * lt, cx);    // NB mass at zero
            double lu, ru;
            fast_log_rcp<CAREFUL>(elam + tphi, cx, lu, ru);
            ld = lu - l1p;
            if (SCORE) {
                const double m = mu * r;
                se = -phi * m * tphi * ru;
                sz = fma(elam, ru, -pi);
                sp = tphi * (lt + m) * phi * ru;           // 1 - t = mu / (mu + phi)
            }
        }
    }
};

// ----- zi.binomial() -- R/Fit_Funs.R:766-827 -----------------------------------
struct FamZiBinomial : FamBase {
    static constexpr bool LIN_ETA = true;
    static constexpr double EMU_LO = 1e-13;      // logit link: |eta| <= 30 - margin
    static constexpr double EMU_HI = 1e13;
    static constexpr double LOG_EMU_LO = -29.93;
    static constexpr double LOG_EMU_HI = 29.93;
    static constexpr bool HAS_ZI = true;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double y, double N, const FamPar&, double& cll, double& cphi) {
        cll = (y == 0.0 || y == N) ? 0.0 : lgamma(N + 1.0) - lgamma(y + 1.0) - lgamma(N - y + 1.0);
        cphi = 0.0;
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double N, GLMMA_POINT_ARGS) {
        double mu, lt, L, pi, l1p;
        link_logit<CAREFUL>(eta, emu, cx, mu, lt, L);
        if (YC != 0) { l1p = 0.0; pi = 0.0; }
        else if (ZT) { l1p = zl1p; pi = zpi; } else zero_part<CAREFUL>(ez, elam, cx, l1p, pi);
        if (YC == 1 || (YC == 0 && y > 0.0)) {
            ld = (CAREFUL ? fma(y, lt, -N * L) : -N * L) - l1p;        // fast variant: + y eta at node level
            if (SCORE) { se = fma(-N, mu, y); sz = -pi; }
        } else {
            const double F = fast_exp(-N * L, cx);         // (1 - mu)^N
            double lu, ru;
            fast_log_rcp<CAREFUL>(elam + F, cx, lu, ru);
            ld = lu - l1p;
            if (SCORE) {
                se = -N * mu * F * ru;
                // lambda / (lambda + F) - pi = pi (1 - F) / (lambda + F); hoisted: the first term alone
                sz = YC != 0 ? elam * ru : pi * (1.0 - F) * ru;
            }
        }
        sp = 0.0;
    }
};

// zero part shared by the hurdle families (e.g. R/Fit_Funs.R:652-655, 670-676)
template <bool CAREFUL, bool ZT = false, int YC = 0>
__device__ __forceinline__ void hurdle_zero(double ez, double elam, const MathCtx& cx, bool pos, double& lz, double& sz,
                                            double zl1p = 0.0, double zpi = 0.0) {
    double pi, l1p;
    if (YC != 0) { l1p = 0.0; pi = 0.0; }
    else if (ZT) { l1p = zl1p; pi = zpi; } else zero_part<CAREFUL>(ez, elam, cx, l1p, pi);
    // positive: log(1 - pi) = -log(1 + e^ez); zero: log pi = ez - log(1 + e^ez)
    lz = pos ? -l1p : ez - l1p;
    sz = pos ? -pi : 1.0 - pi;
}

// ----- hurdle.poisson() -- R/Fit_Funs.R:642-688 --------------------------------
struct FamHurdlePoisson : FamBase {
    static constexpr bool ZERO_FREE = true;
    static constexpr bool Y_TABLE = true;
    static constexpr bool HAS_ZI = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar&, double& cll, double& cphi) {
        cll = (y > 0.0) ? -lgamma(y + 1.0) : 0.0; cphi = 0.0;
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double, GLMMA_POINT_ARGS) {
        const bool pos = YC == 0 ? y > 0.0 : YC == 1;
        double lz;
        hurdle_zero<CAREFUL, ZT, YC>(ez, elam, cx, pos, lz, sz, zl1p, zpi);
        ld = lz; se = 0.0; sp = 0.0;
        if (pos) {
            double mu, lmu;
            link_log<CAREFUL>(eta, emu, mu, lmu);
            if (!CAREFUL || mu < 1e300) {
                // zero-truncated poisson: log f = y eta - mu - log(1 - e^-mu); score = y - mu / (1 - e^-mu)
                // (also the generic kernel's form: mu >= eps by the link clamp, so 1 - e^-mu >= 2e-16 is in the
                // table-based logarithm's range)
                const double om = fast_one_minus_expneg(mu, cx);
                double Lo, ro;
                fast_log_rcp<false>(om, cx, Lo, ro);
                ld += fma(y, eta, -mu) - Lo;
                if (SCORE) se = fma(-mu, ro, y);
            } else {
                const double em1 = expm1(-mu);            // -(1 - e^-mu)
                ld += -mu + y * eta - log(-em1);
                if (SCORE) se = -mu + y + (em1 + 1.0) * mu / em1;
            }
        }
    }
};

// ----- hurdle.negative.binomial() -- R/Fit_Funs.R:690-764 ----------------------
struct FamHurdleNegBin : FamBase {
    static constexpr bool ZERO_FREE = true;
    static constexpr bool Y_TABLE = true;
    static constexpr bool HAS_ZI = true;
    static constexpr bool HAS_PHI = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar& fp, double& cll, double& cphi) {
        if (y > 0.0) nb_const(y, fp, cll, cphi); else { cll = 0.0; cphi = 0.0; }
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double, GLMMA_POINT_ARGS) {
        const bool pos = YC == 0 ? y > 0.0 : YC == 1;
        double lz;
        hurdle_zero<CAREFUL, ZT, YC>(ez, elam, cx, pos, lz, sz, zl1p, zpi);
        ld = lz; se = 0.0; sp = 0.0;
        if (pos) {
            const double phi = fp.a[0];
            double mu, lmu, L, r;
            link_log<CAREFUL>(eta, emu, mu, lmu);
            fast_log_rcp<CAREFUL>(mu + phi, cx, L, r);
            const double yp = y + phi;
            const double lt = fp.a[1] - L;            // log(phi / (phi + mu)) = -log(1 + mu/phi)
            double lom, ratio;                        // log(1 - t^phi) and t^phi / (1 - t^phi), t = phi / (phi + mu)
            if (!CAREFUL || (-phi * lt >= 1e-290 && -phi * lt < 1e300)) {
                const double om = fast_one_minus_expneg(-phi * lt, cx);
                double ro;
                fast_log_rcp<false>(om, cx, lom, ro);
                ratio = (1.0 - om) * ro;
            } else {
                const double em1 = expm1(phi * lt);   // (1 + mu/phi)^-phi - 1
                lom = log(-em1);
                ratio = -(em1 + 1.0) / em1;
            }
            ld += y * lmu - yp * L - lom;
            if (SCORE) {
                const double m = mu * r;
                se = y - yp * m - mu * (phi * r) * ratio;
                sp = phi * (-L - yp * r + ratio * (m + lt));
            }
        }
    }
};

// ----- hurdle.lognormal() -- R/Fit_Funs.R:829-885; a[0] = sigma, a[1] = log sigma,
// a[2] = 1/sigma^2 ---------------------------------------------------------------
struct FamHurdleLogNormal : FamBase {
    static constexpr bool ZERO_FREE = true;
    static constexpr bool HAS_ZI = true;
    static constexpr bool HAS_PHI = true;
    static constexpr bool USES_Y2 = true;
    static constexpr bool NEED_EMU = false;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar& fp, double& cll, double& cphi) {
        if (y > 0.0) { cll = -fp.a[1] - 0.5 * GLMMA_LOG_2PI; cphi = -1.0; }
        else { cll = 0.0; cphi = 0.0; }
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double ly, GLMMA_POINT_ARGS) {
        const bool pos = YC == 0 ? y > 0.0 : YC == 1;
        double lz;
        hurdle_zero<CAREFUL, ZT, YC>(ez, elam, cx, pos, lz, sz, zl1p, zpi);
        ld = lz; se = 0.0; sp = 0.0;
        if (pos) {
            const double d = ly - eta;
            const double d2 = d * d * fp.a[2];
            ld -= 0.5 * d2;
            if (SCORE) { se = d * fp.a[2]; sp = d2; }
        }
    }
};

// beta pieces; a[0] = phi, a[1] = lgamma(phi), a[2] = digamma(phi)
template <bool SCORE, bool CAREFUL>
__device__ __forceinline__ void beta_point(double ly, double l1y, double eta, double emu, const FamPar& fp,
                                           const MathCtx& cx, double& ld, double& se, double& sp) {
    const double phi = fp.a[0];
    double mu, lt, L;
    link_logit<CAREFUL>(eta, emu, cx, mu, lt, L);
    const double a = mu * phi;
    const double b = phi - a;
    const double dl = ly - l1y;
    double lga, lgb, pa, pb;
    if (fmin(a, b) >= 1e-290 && fmax(a, b) <= 1e25) {
        // joint table-based lgamma / digamma (fastmath.cuh); warp-uniform in practice
        fast_lgamma_digamma(a, cx, lga, pa);
        fast_lgamma_digamma(b, cx, lgb, pb);
    } else {
        lga = lgamma(a); lgb = lgamma(b);
        pa = SCORE ? glmma_digamma(a) : 0.0;
        pb = SCORE ? glmma_digamma(b) : 0.0;
    }
    ld = a * dl - lga - lgb;
    if (SCORE) {
        se = phi * (pb - pa + dl) * (mu - mu * mu);
        sp = phi * (mu * (ly - pa) + (1.0 - mu) * (l1y - pb));
    }
}

// ----- beta.fam() -- R/Fit_Funs.R:887-930 ---------------------------------------
struct FamBeta : FamBase {
    static constexpr int GENERIC_UNROLL = 2;
    static constexpr double EMU_LO = 1e-13;      // logit link: |eta| <= 30 - margin
    static constexpr double EMU_HI = 1e13;
    static constexpr double LOG_EMU_LO = -29.93;
    static constexpr double LOG_EMU_HI = 29.93;
    static constexpr bool HAS_PHI = true;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double ly, double l1y, const FamPar& fp, double& cll, double& cphi) {
        cll = fp.a[1] - ly + (fp.a[0] - 1.0) * l1y;
        cphi = fp.a[0] * fp.a[2];
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double ly, double l1y, GLMMA_POINT_ARGS) {
        beta_point<SCORE, CAREFUL>(ly, l1y, eta, emu, fp, cx, ld, se, sp);
        sz = 0.0;
    }
};

// ----- hurdle.beta.fam() -- R/Fit_Funs.R:932-1004 (scores evaluated at the true mu;
// the reference's attr(mu_y) <- eta slip, SURVEY appendix G, is not replicated) ---
struct FamHurdleBeta : FamBase {
    static constexpr bool ZERO_FREE = true;
    __device__ static __forceinline__ bool is_pos(double ly) { return ly > -INFINITY; }
    static constexpr int GENERIC_UNROLL = 2;
    static constexpr double EMU_LO = 1e-13;      // logit link: |eta| <= 30 - margin
    static constexpr double EMU_HI = 1e13;
    static constexpr double LOG_EMU_LO = -29.93;
    static constexpr double LOG_EMU_HI = 29.93;
    static constexpr bool HAS_ZI = true;
    static constexpr bool HAS_PHI = true;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double ly, double l1y, const FamPar& fp, double& cll, double& cphi) {
        if (ly > -INFINITY) { cll = fp.a[1] - ly + (fp.a[0] - 1.0) * l1y; cphi = fp.a[0] * fp.a[2]; }
        else { cll = 0.0; cphi = 0.0; }
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double ly, double l1y, GLMMA_POINT_ARGS) {
        const bool pos = YC == 0 ? ly > -INFINITY : YC == 1;
        double lz;
        hurdle_zero<CAREFUL, ZT, YC>(ez, elam, cx, pos, lz, sz, zl1p, zpi);
        ld = lz; se = 0.0; sp = 0.0;
        if (pos) {
            double l2;
            beta_point<SCORE, CAREFUL>(ly, l1y, eta, emu, fp, cx, l2, se, sp);
            ld += l2;
        }
    }
};

// ----- Gamma.fam() -- R/Fit_Funs.R:1323-1358; a[0] = phi, a[1] = log phi,
// a[2] = lgamma(phi), a[3] = digamma(phi) ----------------------------------------
struct FamGamma : FamBase {
    static constexpr bool HAS_PHI = true;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double, double ly, const FamPar& fp, double& cll, double& cphi) {
        const double phi = fp.a[0];
        cll = (phi - 1.0) * ly - fp.a[2] + phi * fp.a[1];
        cphi = phi * (ly + fp.a[1] + 1.0 - fp.a[3]);
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double, GLMMA_POINT_ARGS) {
        const double phi = fp.a[0];
        double mu, lmu;
        link_log<CAREFUL>(eta, emu, mu, lmu);                   // mu = pmax(exp(eta), eps)
        const double ymu = y / mu;
        ld = -phi * (ymu + lmu);
        if (SCORE) { se = phi * (ymu - 1.0); sp = ld; }
        sz = 0.0;
    }
};

// ----- censored.normal() -- R/Fit_Funs.R:1360-1435; a[0] = sigma, a[1] = log sigma,
// a[2] = 1/sigma ------------------------------------------------------------------
struct FamCensoredNormal : FamBase {
    static constexpr int GENERIC_UNROLL = 2;
    // Register variant: the Gaussian part of the cluster (its uncensored observations) is one quadratic
    // form per node, only the censored observations are evaluated per point (kernels.cuh, FAST_KIND 3)
    static constexpr int FAST_KIND = 3;
    static constexpr bool HAS_PHI = true;
    static constexpr bool USES_Y2 = true;
    static constexpr bool NEED_EMU = false;
    __device__ static __forceinline__ void obs_const(double, double ind, const FamPar& fp, double& cll, double& cphi) {
        if (ind == 0.0) { cll = -fp.a[1] - 0.5 * GLMMA_LOG_2PI; cphi = -1.0; }
        else { cll = 0.0; cphi = 0.0; }
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double ind, GLMMA_POINT_ARGS) {
        const double sigma = fp.a[0], isig = fp.a[2];
        // (the generic kernel takes this form too for every finite argument of moderate size: the library erfc /
        // log path below cost ~600 instructions per censored point -- 24-observation clusters ran 14x slower per
        // point than 16-observation ones)
        if (!CAREFUL || fabs((y - eta) * isig) < 1e100) {
            // Register variant: the same quantities from the table-based Gaussian tail functions
            // (fast_gauss_tail), with x = t (left-censored) or -t (right-censored), B = Phi(x),
            // Bf = max(B, sqrt(eps)) where the reference floors it:
            //   left :  score_eta = -lambda(x) / sigma,                  score_phi = -t phi / max(B, sqrt(eps))
            //   right:  score_eta = (sigma phi - eta (Bf - B)) / (Bf sigma^2),  score_phi = (B - Bf + t phi) / Bf
            // which are the reference's expressions with P = 1 - B substituted (they differ from its
            // floating-point evaluation by rounding noise only, and agree exactly when no floor is active).
            const double t = (y - eta) * isig;
            sz = 0.0;
            if (ind == 0.0) {
                ld = -0.5 * t * t;
                if (SCORE) { se = t * isig; sp = t * t; } else { se = 0.0; sp = 0.0; }
            } else {
                const bool left = ind == 1.0;
                const double x = left ? t : -t;
                double lam, dn, B;
                fast_gauss_tail(x, cx, ld, lam, dn, B);
                const double Bf = fmax(B, GLMMA_SQRT_DBL_EPS);
                if (B >= GLMMA_SQRT_DBL_EPS) {
                    se = (left ? -lam : lam) * isig;
                    sp = -x * lam;
                } else if (left) {
                    se = -lam * isig;
                    sp = -t * dn / Bf;
                } else {
                    se = (sigma * dn - eta * (Bf - B)) / Bf * isig * isig;
                    sp = (B - Bf + t * dn) / Bf;
                }
            }
            return;
        }

        const double t = (y - eta) * isig;
        sz = 0.0; se = 0.0; sp = 0.0;
        if (ind == 0.0) {
            ld = -0.5 * t * t;
            if (SCORE) { se = t * isig; sp = t * t; }
        } else if (ind == 1.0) {
            ld = glmma_log_ndtr(t);
            if (SCORE) {
                const double dn = GLMMA_INV_SQRT_2PI * exp(-0.5 * t * t);
                const double P = glmma_ndtr(t);
                se = -dn / (sigma * P);
                sp = -t * dn / fmax(P, GLMMA_SQRT_DBL_EPS);
            }
        } else {
            ld = glmma_log_ndtr(-t);
            if (SCORE) {
                // the reference's algebraic forms, including the sqrt(eps) floor on 1 - Phi
                const double dn = GLMMA_INV_SQRT_2PI * exp(-0.5 * t * t);
                const double P = glmma_ndtr(t);
                const double B = fmax(glmma_ndtr(-t), GLMMA_SQRT_DBL_EPS);
                const double A = eta * P - sigma * dn;
                se = (-A / B + eta * (1.0 - B) / B) * isig * isig;
                sp = (1.0 - B) / B - (P - t * dn) / B;
            }
        }
    }
};

// ----- students.t(df) -- R/Fit_Funs.R:1006-1041; identity link, df fixed by the family object.
// a[0] = sigma, a[1] = lgamma((df+1)/2) - lgamma(df/2) - log(df pi)/2 - log sigma, a[2] = 1/sigma,
// a[3] = df + 1, a[4] = 1/df, a[5] = (df + 1)/2 --------------------------------------------------
struct FamStudentsT : FamBase {
    static constexpr bool HAS_PHI = true;
    static constexpr bool NEED_EMU = false;
    __device__ static __forceinline__ void obs_const(double, double, const FamPar& fp, double& cll, double& cphi) {
        cll = fp.a[1]; cphi = -1.0;
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double, GLMMA_POINT_ARGS) {
        const double d = y - eta;
        const double t = d * fp.a[2];
        const double w = t * t * fp.a[4];               // (y - mu)^2 / (df sigma^2)
        double L, r;
        fast_log_rcp<CAREFUL>(1.0 + w, cx, L, r);
        ld = -fp.a[5] * L;
        if (SCORE) {
            se = d * fp.a[2] * fp.a[2] * fp.a[3] * fp.a[4] * r;
            sp = fp.a[3] * w * r;
        } else { se = 0.0; sp = 0.0; }
        sz = 0.0;
    }
};

// ----- beta.binomial(), logit -- R/Fit_Funs.R:1245-1321.  y1 = successes, y2 = trials N.
// a[0] = phi, a[1] = lgamma(phi), a[2] = digamma(phi).  With A = phi mu, B = phi (1 - mu):
//   log f = lchoose(N, y) + [lgamma(y + A) - lgamma(A)] + [lgamma(N - y + B) - lgamma(B)] + lgamma(phi) - lgamma(N + phi)
// For integer counts the bracketed differences are finite sums  sum_{i < y} log(A + i)  and their
// derivatives  sum_{i < y} 1 / (A + i), i.e. one fast_log_rcp per unit of y (Bernoulli data: one per
// point); counts above GLMMA_BB_SUM_MAX (or non-integer responses) take lgamma / digamma. -----------
#define GLMMA_BB_SUM_MAX 24
// (the lgamma / digamma path is out of line: inlined at every point of every unrolled observation it made this
// family's translation unit the longest to compile and the largest in the library)
static __device__ __noinline__ void bb_rising_libm(double x, double cnt, double* out) {
    out[0] = lgamma(x + cnt) - lgamma(x);
    out[1] = glmma_digamma(x + cnt) - glmma_digamma(x);
}
__device__ __forceinline__ void bb_rising(double x, double cnt, const MathCtx& cx, double& dlg, double& dpsi) {
    // lgamma(x + cnt) - lgamma(x) and digamma(x + cnt) - digamma(x), cnt >= 0
    const int n = (int)cnt;
    if ((double)n == cnt && n <= GLMMA_BB_SUM_MAX && x >= 1e-290 && x <= 1e290) {
        dlg = 0.0; dpsi = 0.0;
        for (int i = 0; i < n; ++i) {
            double L, r;
            fast_log_rcp<false>(x + (double)i, cx, L, r);       // in range by the test above: no library fallback inlined
            dlg += L; dpsi += r;
        }
    } else {
        double o[2];
        bb_rising_libm(x, cnt, o);
        dlg = o[0]; dpsi = o[1];
    }
}
template <int LINK>   // 1 logit, 5 cloglog (enum glmma_link)
struct FamBetaBinomialL : FamBase {
    static constexpr int GENERIC_UNROLL = 2;
    // logit link: |eta| <= 30 - margin; cloglog: eps <= mu <= 1 - eps  <=>  1e-15 <= exp(eta) <= 35 (as binomial)
    static constexpr double EMU_LO = LINK == 5 ? 1e-15 : 1e-13;
    static constexpr double EMU_HI = LINK == 5 ? 35.0 : 1e13;
    static constexpr double LOG_EMU_LO = LINK == 5 ? -34.5 : -29.93;
    static constexpr double LOG_EMU_HI = LINK == 5 ? 3.55 : 29.93;
    static constexpr bool HAS_PHI = true;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double y, double N, const FamPar& fp, double& cll, double& cphi) {
        const double lch = (y == 0.0 || y == N) ? 0.0 : lgamma(N + 1.0) - lgamma(y + 1.0) - lgamma(N - y + 1.0);
        cll = lch + fp.a[1] - lgamma(N + fp.a[0]);
        cphi = fp.a[0] * (fp.a[2] - glmma_digamma(N + fp.a[0]));
    }
    template <bool SCORE, bool CAREFUL, bool ZT = false, int YC = 0>
    __device__ static __forceinline__ void point(double y, double N, GLMMA_POINT_ARGS) {
        const double phi = fp.a[0];
        double mu, dmu;
        if (LINK == 5) {
            // make.link("cloglog")$linkinv with its clamps; mu.eta is the reference's own expression
            // -(1 - mu) log(1 - mu) on the ROUNDED 1 - mu (R/Fit_Funs.R:1288-1290), so that both sides are ill
            // conditioned in the same way where mu is tiny
            if (CAREFUL) mu = fmax(fmin(-expm1(-exp(eta)), 1.0 - GLMMA_DBL_EPS), GLMMA_DBL_EPS);
            else mu = fast_one_minus_expneg(emu, cx);
            const double q = 1.0 - mu;
            double Lq, rq;
            if (CAREFUL) { Lq = log(q); rq = 0.0; } else fast_log_rcp<false>(q, cx, Lq, rq);
            dmu = -q * Lq;
        } else {
            double lt, L;
            link_logit<CAREFUL>(eta, emu, cx, mu, lt, L);
            dmu = mu - mu * mu;
        }
        const double A = phi * mu, B = LINK == 5 ? phi * (1.0 - mu) : phi - A;
        double lgA, psA, lgB, psB;
        bb_rising(A, y, cx, lgA, psA);
        bb_rising(B, N - y, cx, lgB, psB);
        ld = lgA + lgB;
        if (SCORE) {
            se = phi * (psA - psB) * dmu;
            sp = phi * (mu * psA + (1.0 - mu) * psB);
        } else { se = 0.0; sp = 0.0; }
        sz = 0.0;
    }
};
typedef FamBetaBinomialL<1> FamBetaBinomial;
