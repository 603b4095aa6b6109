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

// ----- link pieces with R's clamps (stats::make.link) ------
// log link: mu = pmax(exp(eta), eps); log mu follows the clamp
__device__ __forceinline__ void link_log(double eta, double& mu, double& lmu) {
    const bool lo = eta < GLMMA_LOG_DBL_EPS;
    mu = lo ? GLMMA_DBL_EPS : exp(eta);
    lmu = lo ? GLMMA_LOG_DBL_EPS : eta;
}
// logit link: t = exp(eta) clamped to [eps, 1/eps] outside +-30; mu = t / (1 + t).
// Returns mu, log(mu) and log(1 - mu) evaluated without cancellation.
__device__ __forceinline__ void link_logit(double eta, double& mu, double& lp, double& lq) {
    double t, lt;
    if (eta < -30.0) { t = GLMMA_DBL_EPS; lt = GLMMA_LOG_DBL_EPS; }
    else if (eta > 30.0) { t = GLMMA_INV_DBL_EPS; lt = -GLMMA_LOG_DBL_EPS; }
    else { t = exp(eta); lt = eta; }
    const double u = 1.0 + t;
    mu = t / u;
    const double L = (t < 0.5) ? log1p(t) : log(u);
    lq = -L;
    lp = lt - L;
}

struct FamBase {
    static constexpr bool HAS_ZI = false;    // family has a zero part (eta_zi)
    static constexpr bool HAS_PHI = false;   // family has a dispersion parameter
    static constexpr bool USES_Y2 = false;
    __device__ static __forceinline__ void obs_const(double, double, const FamPar&, double& cll, double& cphi) {
        cll = 0.0; cphi = 0.0;
    }
};

// ----- stats::binomial(), logit (canonical) -- R/Fit_Funs.R:429-438, :130-149 --
struct FamBinomialLogit : FamBase {
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double y, double N, const FamPar&, double& cll, double& cphi) {
        cll = (y == 0.0 || y == N) ? 0.0 : lgamma(N + 1.0) - lgamma(y + 1.0) - lgamma(N - y + 1.0);
        cphi = 0.0;
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double N, double eta, double, const FamPar&,
                                                 double& ld, double& se, double& sz, double& sp) {
        double mu, lp, lq;
        link_logit(eta, mu, lp, lq);
        // y log p + (N - y) log q; 0 * (-inf) cannot occur (lp, lq are finite)
        ld = y * lp + (N - y) * lq;
        if (SCORE) se = y - N * mu;
        sz = 0.0; sp = 0.0;
    }
};

// ----- stats::binomial(), probit / cloglog -- score via (y - N mu) mu'(eta) / V(mu),
// R/Fit_Funs.R:150-167 ---------------------------------------------------------
template <int LINK>   // 4 probit, 5 cloglog (enum glmma_link)
struct FamBinomialOther : FamBase {
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double y, double N, const FamPar& fp, double& cll, double& cphi) {
        FamBinomialLogit::obs_const(y, N, fp, cll, cphi);
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double N, double eta, double, const FamPar&,
                                                 double& ld, double& se, double& sz, double& sp) {
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
        sz = 0.0; sp = 0.0;
    }
};

// ----- stats::poisson(), log -- R/Fit_Funs.R:440-445 ---------------------------
struct FamPoisson : FamBase {
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar&, double& cll, double& cphi) {
        cll = -lgamma(y + 1.0); cphi = 0.0;
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double, double eta, double, const FamPar&,
                                                 double& ld, double& se, double& sz, double& sp) {
        double mu, lmu;
        link_log(eta, mu, lmu);
        ld = y * lmu - mu;
        if (SCORE) se = y - mu;
        sz = 0.0; sp = 0.0;
    }
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
    static constexpr bool HAS_PHI = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar& fp, double& cll, double& cphi) {
        nb_const(y, fp, cll, cphi);
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double, double eta, double, const FamPar& fp,
                                                 double& ld, double& se, double& sz, double& sp) {
        const double phi = fp.a[0];
        double mu, lmu;
        link_log(eta, mu, lmu);
        const double t = mu + phi;
        const double L = log(t);
        const double yp = y + phi;
        ld = y * lmu - yp * L;
        if (SCORE) {
            const double r = 1.0 / t;
            se = y - yp * (mu * r);
            sp = -phi * (L + yp * r);
        }
        sz = 0.0;
    }
};

// ----- zi.poisson() -- R/Fit_Funs.R:515-563 ------------------------------------
struct FamZiPoisson : FamBase {
    static constexpr bool HAS_ZI = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar&, double& cll, double& cphi) {
        cll = (y > 0.0) ? -lgamma(y + 1.0) : 0.0; cphi = 0.0;
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double, double eta, double ez, const FamPar&,
                                                 double& ld, double& se, double& sz, double& sp) {
        double mu, lmu;
        link_log(eta, mu, lmu);
        double pi, l1p;                    // plogis(eta_zi), log(1 + lambda)
        glmma_sigmoid_l1pe(ez, pi, l1p);
        if (y > 0.0) {
            ld = y * lmu - mu - l1p;
            if (SCORE) { se = y - mu; sz = -pi; }
        } else {
            // log(lambda + exp(-mu)) - log(1 + lambda); lambda capped where it would overflow
            const double lam = exp(fmin(ez, 700.0));
            const double E = exp(-mu);
            const double u = lam + E;
            ld = log(u) - l1p;
            if (SCORE) {
                const double ru = 1.0 / u;
                se = -mu * E * ru;
                sz = lam * ru - pi;
            }
        }
        sp = 0.0;
    }
};

// ----- zi.negative.binomial() -- R/Fit_Funs.R:565-640 --------------------------
struct FamZiNegBin : FamBase {
    static constexpr bool HAS_ZI = true;
    static constexpr bool HAS_PHI = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar& fp, double& cll, double& cphi) {
        if (y > 0.0) nb_const(y, fp, cll, cphi); else { cll = 0.0; cphi = 0.0; }
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double, double eta, double ez, const FamPar& fp,
                                                 double& ld, double& se, double& sz, double& sp) {
        const double phi = fp.a[0];
        double mu, lmu;
        link_log(eta, mu, lmu);
        const double t = mu + phi;
        const double L = log(t);
        double pi, l1p;
        glmma_sigmoid_l1pe(ez, pi, l1p);
        if (y > 0.0) {
            const double yp = y + phi;
            ld = y * lmu - yp * L - l1p;
            if (SCORE) {
                const double r = 1.0 / t;
                se = y - yp * (mu * r);
                sp = -phi * (L + yp * r);
                sz = -pi;
            }
        } else {
            const double lt = fp.a[1] - L;            // log(phi / (phi + mu))
            const double tphi = exp(phi * lt);        // NB mass at zero
            const double lam = exp(fmin(ez, 700.0));
            const double u = lam + tphi;
            ld = log(u) - l1p;
            if (SCORE) {
                const double ru = 1.0 / u;
                const double m = mu / t;
                se = -phi * m * tphi * ru;
                sz = lam * ru - pi;
                sp = tphi * (lt + m) * phi * ru;      // 1 - t = mu / (mu + phi)
            }
        }
    }
};

// ----- zi.binomial() -- R/Fit_Funs.R:766-827 -----------------------------------
struct FamZiBinomial : FamBase {
    static constexpr bool HAS_ZI = true;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double y, double N, const FamPar&, double& cll, double& cphi) {
        cll = (y == 0.0 || y == N) ? 0.0 : lgamma(N + 1.0) - lgamma(y + 1.0) - lgamma(N - y + 1.0);
        cphi = 0.0;
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double N, double eta, double ez, const FamPar&,
                                                 double& ld, double& se, double& sz, double& sp) {
        double mu, lp, lq;
        link_logit(eta, mu, lp, lq);
        double pi, l1p;
        glmma_sigmoid_l1pe(ez, pi, l1p);
        if (y > 0.0) {
            ld = y * lp + (N - y) * lq - l1p;
            if (SCORE) { se = y - N * mu; sz = -pi; }
        } else {
            const double F = exp(N * lq);             // (1 - mu)^N
            const double lam = exp(fmin(ez, 700.0));
            const double u = lam + F;
            ld = log(u) - l1p;
            if (SCORE) {
                const double ru = 1.0 / u;
                se = -N * mu * F * ru;
                sz = pi * (1.0 - F) * ru;
            }
        }
        sp = 0.0;
    }
};

// zero part shared by the hurdle families (e.g. R/Fit_Funs.R:652-655, 670-676)
__device__ __forceinline__ void hurdle_zero(double ez, bool pos, double& lz, double& sz) {
    double pi, l1p;
    glmma_sigmoid_l1pe(ez, pi, l1p);
    // positive: log(1 - pi) = -log(1 + e^ez); zero: log pi = ez - log(1 + e^ez)
    lz = pos ? -l1p : ez - l1p;
    sz = pos ? -pi : 1.0 - pi;
}

// ----- hurdle.poisson() -- R/Fit_Funs.R:642-688 --------------------------------
struct FamHurdlePoisson : FamBase {
    static constexpr bool HAS_ZI = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar&, double& cll, double& cphi) {
        cll = (y > 0.0) ? -lgamma(y + 1.0) : 0.0; cphi = 0.0;
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double, double eta, double ez, const FamPar&,
                                                 double& ld, double& se, double& sz, double& sp) {
        const bool pos = y > 0.0;
        double lz;
        hurdle_zero(ez, pos, lz, sz);
        ld = lz; se = 0.0; sp = 0.0;
        if (pos) {
            double mu, lmu;
            link_log(eta, mu, lmu);
            const double em1 = expm1(-mu);            // -(1 - e^-mu)
            ld += -mu + y * eta - log(-em1);
            if (SCORE) se = -mu + y + (em1 + 1.0) * mu / em1;
        }
    }
};

// ----- hurdle.negative.binomial() -- R/Fit_Funs.R:690-764 ----------------------
struct FamHurdleNegBin : FamBase {
    static constexpr bool HAS_ZI = true;
    static constexpr bool HAS_PHI = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar& fp, double& cll, double& cphi) {
        if (y > 0.0) {
            nb_const(y, fp, cll, cphi);
            // the reference's hurdle NB score has digamma(y+phi) - digamma(phi) + log phi + 1
            // in common with nb_const: nothing to adjust
        } else { cll = 0.0; cphi = 0.0; }
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double, double eta, double ez, const FamPar& fp,
                                                 double& ld, double& se, double& sz, double& sp) {
        const bool pos = y > 0.0;
        double lz;
        hurdle_zero(ez, pos, lz, sz);
        ld = lz; se = 0.0; sp = 0.0;
        if (pos) {
            const double phi = fp.a[0];
            double mu, lmu;
            link_log(eta, mu, lmu);
            const double t = mu + phi;
            const double L = log(t);
            const double yp = y + phi;
            const double lt = fp.a[1] - L;            // log(phi / (phi + mu)) = -log(1 + mu/phi)
            const double tphi = exp(phi * lt);        // (1 + mu/phi)^-phi
            const double om = -expm1(phi * lt);       // 1 - (1 + mu/phi)^-phi
            ld += y * lmu - yp * L - log(om);
            if (SCORE) {
                const double r = 1.0 / t;
                const double m = mu * r;
                const double ratio = tphi / om;
                se = y - yp * m - mu * (phi * r) * ratio;
                sp = phi * (-L - yp * r + ratio * (m + lt));
            }
        }
    }
};

// ----- hurdle.lognormal() -- R/Fit_Funs.R:829-885; a[0] = sigma, a[1] = log sigma,
// a[2] = 1/sigma^2 ---------------------------------------------------------------
struct FamHurdleLogNormal : FamBase {
    static constexpr bool HAS_ZI = true;
    static constexpr bool HAS_PHI = true;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double y, double, const FamPar& fp, double& cll, double& cphi) {
        if (y > 0.0) { cll = -fp.a[1] - 0.5 * GLMMA_LOG_2PI; cphi = -1.0; }
        else { cll = 0.0; cphi = 0.0; }
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double ly, double eta, double ez, const FamPar& fp,
                                                 double& ld, double& se, double& sz, double& sp) {
        const bool pos = y > 0.0;
        double lz;
        hurdle_zero(ez, pos, lz, sz);
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
template <bool SCORE>
__device__ __forceinline__ void beta_point(double ly, double l1y, double eta, const FamPar& fp,
                                           double& ld, double& se, double& sp) {
    const double phi = fp.a[0];
    double mu, lp, lq;
    link_logit(eta, mu, lp, lq);
    const double a = mu * phi;
    const double b = phi - a;
    const double dl = ly - l1y;
    ld = a * dl - lgamma(a) - lgamma(b);
    if (SCORE) {
        const double pa = glmma_digamma(a);
        const double pb = glmma_digamma(b);
        se = phi * (pb - pa + dl) * (mu - mu * mu);
        sp = phi * (mu * (ly - pa) + (1.0 - mu) * (l1y - pb));
    }
}

// ----- beta.fam() -- R/Fit_Funs.R:887-930 ---------------------------------------
struct FamBeta : FamBase {
    static constexpr bool HAS_PHI = true;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double ly, double l1y, const FamPar& fp, double& cll, double& cphi) {
        cll = fp.a[1] - ly + (fp.a[0] - 1.0) * l1y;
        cphi = fp.a[0] * fp.a[2];
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double ly, double l1y, double eta, double, const FamPar& fp,
                                                 double& ld, double& se, double& sz, double& sp) {
        beta_point<SCORE>(ly, l1y, eta, fp, ld, se, sp);
        sz = 0.0;
    }
};

// ----- hurdle.beta.fam() -- R/Fit_Funs.R:932-1004 (scores evaluated at the true mu;
// the reference's attr(mu_y) <- eta slip, SURVEY appendix G, is not replicated) ---
struct FamHurdleBeta : FamBase {
    static constexpr bool HAS_ZI = true;
    static constexpr bool HAS_PHI = true;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double ly, double l1y, const FamPar& fp, double& cll, double& cphi) {
        if (ly > -INFINITY) { cll = fp.a[1] - ly + (fp.a[0] - 1.0) * l1y; cphi = fp.a[0] * fp.a[2]; }
        else { cll = 0.0; cphi = 0.0; }
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double ly, double l1y, double eta, double ez, const FamPar& fp,
                                                 double& ld, double& se, double& sz, double& sp) {
        const bool pos = ly > -INFINITY;
        double lz;
        hurdle_zero(ez, pos, lz, sz);
        ld = lz; se = 0.0; sp = 0.0;
        if (pos) {
            double l2;
            beta_point<SCORE>(ly, l1y, eta, fp, l2, se, sp);
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
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double, double eta, double, const FamPar& fp,
                                                 double& ld, double& se, double& sz, double& sp) {
        const double phi = fp.a[0];
        const bool lo = eta < GLMMA_LOG_DBL_EPS;           // mu = pmax(exp(eta), eps)
        const double lmu = lo ? GLMMA_LOG_DBL_EPS : eta;
        const double ymu = y * (lo ? GLMMA_INV_DBL_EPS : exp(-eta));
        ld = -phi * (ymu + lmu);
        if (SCORE) { se = phi * (ymu - 1.0); sp = ld; }
        sz = 0.0;
    }
};

// ----- censored.normal() -- R/Fit_Funs.R:1360-1435; a[0] = sigma, a[1] = log sigma,
// a[2] = 1/sigma ------------------------------------------------------------------
struct FamCensoredNormal : FamBase {
    static constexpr bool HAS_PHI = true;
    static constexpr bool USES_Y2 = true;
    __device__ static __forceinline__ void obs_const(double, double ind, const FamPar& fp, double& cll, double& cphi) {
        if (ind == 0.0) { cll = -fp.a[1] - 0.5 * GLMMA_LOG_2PI; cphi = -1.0; }
        else { cll = 0.0; cphi = 0.0; }
    }
    template <bool SCORE>
    __device__ static __forceinline__ void point(double y, double ind, double eta, double, const FamPar& fp,
                                                 double& ld, double& se, double& sz, double& sp) {
        const double sigma = fp.a[0], isig = fp.a[2];
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
