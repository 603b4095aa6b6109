// Shared device/host definitions of the AGQ engine (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define GLMMA_QMAX 4          // max random-effect dimension q = q_y + q_zi compiled in
#define GLMMA_NAGQ_MAX 32     // max nodes per dimension
#define GLMMA_PMAX 64         // max ncol(X)
#define GLMMA_PZIMAX 32       // max ncol(X_zi)
#define GLMMA_NPACK (GLMMA_QMAX * (GLMMA_QMAX + 1) / 2)
// per-warp / per-block partial sums of one evaluation:
//   0 loglik, 1 sum_w, 2 n_nonfinite, 3 g_phi, 4.. S packed upper (d <= e, row-major)
#define GLMMA_NACC (4 + GLMMA_NPACK)
// fused mode: a block partial also carries X'zbar and X_zi'zbar_zi, column l in slot GLMMA_NACC + l
#define GLMMA_FOLD_COLS 32
#define GLMMA_PSTRIDE (GLMMA_NACC + GLMMA_FOLD_COLS)     // doubles per block partial (every mode)

// Cluster-sharded evaluation over several processes (one per GPU): the ranks' result vectors are
// exchanged through peer memory by the finalising block itself (kernels.cuh: exchange_block).
#define GLMMA_MAX_RANKS 16
#define GLMMA_XSLOT 128        // doubles per (parity, rank) slot: the result vector, flag in the last one
struct XchgPar {
    double* peer[GLMMA_MAX_RANKS];   // every rank's exchange buffer as mapped into THIS process (peer[rank]: our own)
    int rank, nranks;                // nranks <= 1: no exchange
    unsigned long long seq;          // evaluation counter, identical on all ranks (they evaluate in lockstep)
};

// zero-copy delivery of the result vector into pinned host memory (kernels.cuh: deliver_host)
struct HostOut {
    double* vec;                          // device-side address of the pinned host vector (null: not used)
    unsigned long long* flag;             // raised to `seq` after the vector is complete
    unsigned long long seq;
};

#define GLMMA_DBL_EPS 2.220446049250313e-16
#define GLMMA_INV_DBL_EPS 4503599627370496.0
#define GLMMA_LOG_DBL_EPS (-36.04365338911715)   // log(.Machine$double.eps)
#define GLMMA_SQRT_DBL_EPS 1.4901161193847656e-08
#define GLMMA_LOG_2PI 1.8378770664093453
#define GLMMA_INV_SQRT_2PI 0.3989422804014327
#define GLMMA_SQRT2 1.4142135623730951
#define GLMMA_INV_SQRT2 0.7071067811865476

// family-level constants derived from phis once per evaluation on the host
struct FamPar {
    double a[6];
};

// fixed effects (prep / score-reduce kernels only)
struct CoefPar {
    double betas[GLMMA_PMAX];
    double gammas[GLMMA_PZIMAX];
};

// what the per-node part of an evaluation needs from theta
struct PriorPar {
    double Dinv[GLMMA_QMAX * GLMMA_QMAX];   // symmetric
    double logprior_const;                  // -0.5 * (q log 2pi + log det D)
};

struct GridPar {
    double x[GLMMA_NAGQ_MAX];     // gauher nodes (descending, as the reference)
    double lw[GLMMA_NAGQ_MAX];    // log(w_a) + x_a^2   (log of one factor of wGH)
    double lw_const;              // (q/2) log 2
    int nagq;
    int K;                        // nagq^q
};

// Device-resident model data (pointers into HBM).  All observation arrays keep
// R's column-major layout: column c of an N x C matrix starts at base + c*N.
struct DevData {
    int64_t N;            // observations
    int64_t n;            // clusters
    int p, q_y, p_zi, q_zi, n_phis;
    int max_ni;           // largest cluster
    const double* y1;     // N  (family-specific transform of y[,1], see engine.cu)
    const double* y2;     // N  (trials / indicator / log terms) or null
    const double* X;      // N x p
    const double* Z;      // N x q_y
    const double* X_zi;   // N x p_zi or null
    const double* Z_zi;   // N x q_zi or null
    const double* offset; // N or null
    const double* offset_zi;
    const double* weights;   // n or null
    const int32_t* row_ptr;  // n + 1
    // theta-dependent per-observation work arrays (prep kernel)
    double* xb;           // N : X beta + offset
    double* xg;           // N : X_zi gamma + offset_zi (zero-part families)
    double* cll;          // N : node-independent part of log f(y_ij)
    double* cphi;         // N : node-independent part of d log f / d phis
    // grid
    double* modes;        // n x q   (cluster-major)
    double* rinv;         // n x q(q+1)/2, packed upper R^-1, row-major over d <= e
    double* chol;         // n x q(q+1)/2, packed upper R in R's U[upper.tri(U,TRUE)] order
    double* logdet;       // n
    // outputs of one evaluation
    double* zbar;         // N : w_i sum_k pi_ik s_eta(ijk)
    double* zbar_zi;      // N
    double* sphi_obs;     // N (contrib only) or null
    double* ll_i;         // n (contrib only) or null
    double* S_i;          // n x q*q (contrib only) or null
    double* csum;         // n x GLMMA_NCSUM per-cluster scalars (cluster_sums_kernel)
    double* clrec;        // n x (q + q(q+1)/2 + 1 + q_y + 1): packed {modes, R^-1, logdet, Z'y1, weight} (register variant)
    int32_t* defer;       // n : 1 = cluster left to the generic eval kernel by the register variant
};

// ------------------------------------------------------------------------------
// math helpers
// ------------------------------------------------------------------------------
__host__ __device__ __forceinline__ double glmma_digamma(double x) {
    // x > 0.  psi(x) = psi(x + m) - sum_{i<m} 1/(x+i): the sum of reciprocals is
    // P'(x)/P(x) for P = prod (x+i), so the shift costs one division; then the
    // asymptotic series at x >= 10 (truncation error < 2e-17 relative).
    double r = 0.0;
    if (x < 10.0) {
        double P = 1.0, dP = 0.0;
#pragma unroll 1
        while (x < 10.0) {
            dP = dP * x + P;
            P *= x;
            x += 1.0;
        }
        r = -dP / P;
    }
    const double f = 1.0 / (x * x);
    double t = 1.0 / 12.0;   // B14/14
    t = 691.0 / 32760.0 - f * t;
    t = 1.0 / 132.0 - f * t;
    t = 1.0 / 240.0 - f * t;
    t = 1.0 / 252.0 - f * t;
    t = 1.0 / 120.0 - f * t;
    t = 1.0 / 12.0 - f * t;
    return r + log(x) - 0.5 / x - f * t;
}

// log(1 + exp(x)), stable
__device__ __forceinline__ double glmma_log1pexp(double x) {
    const double e = exp(-fabs(x));
    return fmax(x, 0.0) + log1p(e);
}

// plogis(x) and log(1 + exp(x)) from one exp
__device__ __forceinline__ void glmma_sigmoid_l1pe(double x, double& sig, double& l1pe) {
    const double e = exp(-fabs(x));
    const double inv = 1.0 / (1.0 + e);
    sig = (x >= 0.0) ? inv : e * inv;
    l1pe = fmax(x, 0.0) + log1p(e);
}

__device__ __forceinline__ double glmma_ndtr(double x) {
    return 0.5 * erfc(-x * GLMMA_INV_SQRT2);
}

// log Phi(x) with accurate tails (R: pnorm(log.p = TRUE))
__device__ __forceinline__ double glmma_log_ndtr(double x) {
    if (x < -1.0) {
        // Phi(x) = 0.5 erfcx(-x/sqrt2) exp(-x^2/2)
        return log(0.5 * erfcx(-x * GLMMA_INV_SQRT2)) - 0.5 * x * x;
    }
    return log1p(-0.5 * erfc(x * GLMMA_INV_SQRT2));
}

// group (sub-warp) reductions; G is a power of two <= 32, gmask names the group's lanes
template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned gmask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ double group_max(double v, unsigned gmask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(gmask, v, o));
    return v;
}
