// fp64 special functions tailored to the AGQ inner loop.
//
// CUDA's exp()/log()/division cost ~35/~40/~15 instructions each on sm_100a, a
// large part of them moves of polynomial coefficients and special-case
// branches (ncu, profiles/r1_*: 240 instructions per 32 quadrature points, of
// which 73 on the FP64 pipe).  The functions below work from two small
// shared-memory tables (fastmath_tables.h) and constant-bank coefficients:
//
//   fast_log_rcp(t)  -> log t and 1/t together   (15 FP64 instructions)
//   fast_exp(x)      -> exp x                    (10 FP64 instructions)
//
// Accuracy (tests/test_fastmath.py, against mpmath): |error| <= 2 ulp for exp
// and 1/t, <= 2 ulp or 2e-16 absolute for log.  Arguments outside the fast
// range (t outside [1e-300, 1e300], |x| > 700, NaN) take the CUDA library path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define GLMMA_HD __host__ __device__ __forceinline__
#else
#define GLMMA_HD inline
#endif

#ifdef __CUDACC__
#define GLMMA_NOINLINE __host__ __device__ __noinline__
#else
#define GLMMA_NOINLINE
#endif
// out-of-range arguments: kept out of line so that the hot loops stay small
static GLMMA_NOINLINE double fm_slow_exp(double x) { return exp(x); }
static GLMMA_NOINLINE double fm_slow_log(double x) { return log(x); }
static GLMMA_NOINLINE double fm_slow_rcp(double x) { return 1.0 / x; }

// Tables + the polynomial / range-reduction constants that are not encodable as
// instruction immediates.  The kernels read the constants through a shared-memory
// pointer: as literals ptxas re-materialises each of them with two moves at every use
// (ncu: 24 of 134 instructions per 32 points were such moves); as loads it keeps the
// hot ones in registers and re-reads the others with one LDS.
//
// On the device both tables are REPLICATED in shared memory so that the data-dependent
// lookups are bank-conflict free: entry i of the log table has 8 copies, lane l reads copy
// l & 7 (a 16-byte access is served per quarter warp: 8 lanes x 16 B = all 32 banks once);
// entry j of the exp table has 16 copies, lane l reads copy l & 15.  With a single copy the
// random indices made the shared-memory pipe the limiter of the whole kernel (ncu r1:
// l1tex data-pipe wavefronts 76 %, 48 M bank conflicts per 200 k clusters).
#ifdef __CUDA_ARCH__
#define FM_LOG_STRIDE 16      // doubles between consecutive entries (8 copies x 2 doubles)
#define FM_EXP_STRIDE 16      // 16 copies x 1 double
#else
#define FM_LOG_STRIDE 2
#define FM_EXP_STRIDE 1
#endif
#define GLMMA_FM_LOG_DOUBLES (128 * 16)
#define GLMMA_FM_EXP_DOUBLES (64 * 16)
#define GLMMA_FM_NCONST 14
struct MathCtx {
    const double* logtab;   // device: replicated table + 2 * (lane & 7); host: 128 x (invc, logc)
    const double* exptab;   // device: replicated table + (lane & 15);    host: 64 x 2^(j/64)
    double c[GLMMA_FM_NCONST];   // constants, order of fastmath_constants(); on the device they come
                                 // from the kernel parameters (constant bank), not from shared memory
    // device: the log table's shared-window address split into its block-uniform base and this lane's
    // byte offset (16 * (lane & 7)), so that a lookup is shift + one 3-input logic op (index bits | lane
    // bits) + a load with a uniform base register
    unsigned lt_base, lt_lane;
};
#define FMC_L6 0      // -1/6
#define FMC_L5 1      // 1/5
#define FMC_L3 2      // 1/3
#define FMC_LN2HI 3
#define FMC_LN2LO 4
#define FMC_E64LN2 5  // 64 / ln2
#define FMC_EMAGIC 6  // 1.5 * 2^52
#define FMC_EHI 7     // -ln2/64 high part
#define FMC_ELO 8     // -ln2/64 low part
#define FMC_E120 9
#define FMC_E24 10
#define FMC_E6 11
#define FMC_LN2 12    // ln 2 rounded to double
#define FMC_LN2S 13   // ln 2 * 2^-20 (stage-wise log: one fma for k ln2 + logc from k * 2^20)
GLMMA_HD MathCtx make_mathctx(const double* logtab, const double* exptab, const double* c) {
    MathCtx m;
    m.logtab = logtab; m.exptab = exptab;
    for (int i = 0; i < GLMMA_FM_NCONST; ++i) m.c[i] = c[i];
#ifdef __CUDA_ARCH__
    m.lt_lane = (threadIdx.x & 7u) * 16u;
    m.lt_base = (unsigned)__cvta_generic_to_shared(logtab) - m.lt_lane;
#else
    m.lt_base = m.lt_lane = 0;
#endif
    return m;
}

GLMMA_HD int fm_hi(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
GLMMA_HD int fm_lo(double x) {
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffff);
#endif
}
GLMMA_HD double fm_make(int hi, int lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x;
#endif
}

#define FM_LN2_HI 0x1.62e42fefa3800p-1     // ln 2 with 11 trailing zero bits: k * LN2_HI is exact
#define FM_LN2_LO 0x1.ef35793c76730p-45
#define FM_64_LN2 0x1.71547652b82fep+6     // 64 / ln 2
#define FM_LN2_64_HI 0x1.62e42fefa0000p-7  // ln2/64, trailing zeros: n * hi exact for |n| < 2^17
#define FM_LN2_64_LO 0x1.cf79abc9e3b3ap-46
#define FM_MAGIC 0x1.8p52
inline void fastmath_constants(double* c) {
    c[0] = -1.0 / 6.0; c[1] = 0.2; c[2] = 1.0 / 3.0; c[3] = FM_LN2_HI; c[4] = FM_LN2_LO;
    c[5] = FM_64_LN2; c[6] = FM_MAGIC; c[7] = -FM_LN2_64_HI; c[8] = -FM_LN2_64_LO;
    c[9] = 1.0 / 120.0; c[10] = 1.0 / 24.0; c[11] = 1.0 / 6.0;
    c[12] = 0x1.62e42fefa39efp-1;
    c[13] = 0x1.62e42fefa39efp-21;
}

// log t and 1/t for t > 0.
//   t = 2^k z, z in [0.6875, 1.375); table entry (invc, logc) of z's sub-interval;
//   r = z invc - 1 (one fma, exact to rounding), |r| < 0.0039
//   log t = k ln2 + logc + r - r^2/2 + r^3/3 - r^4/4 + r^5/5 - r^6/6
//   1/z   = invc (1 - r)(1 + r^2 + r^4 + r^6)
// CHECK = false: the caller guarantees 1e-300 <= t <= 1e300 (no branch, no library fallback)
template <bool CHECK = true>
GLMMA_HD void fast_log_rcp(double t, const MathCtx& cx, double& L, double& rcp) {
    const int hi = fm_hi(t);
    if (CHECK && (unsigned)(hi - 0x01a00000) >= (unsigned)(0x7e300000 - 0x01a00000)) {
        L = fm_slow_log(t);
        rcp = fm_slow_rcp(t);
        return;
    }
    const int tmp = hi - 0x3FE60000;
    const int i = (tmp >> 13) & 127;
    const int k = tmp >> 20;
    const int ke = tmp & 0xfff00000;
    const double z = fm_make(hi - ke, fm_lo(t));
#ifdef __CUDA_ARCH__
    const double2 tc = *reinterpret_cast<const double2*>(cx.logtab + FM_LOG_STRIDE * i);
    const double invc = tc.x, logc = tc.y;
#else
    const double invc = cx.logtab[FM_LOG_STRIDE * i], logc = cx.logtab[FM_LOG_STRIDE * i + 1];
#endif
    const double kd = (double)k;
    const double r = fma(z, invc, -1.0);
    const double r2 = r * r;
    double q = fma(r, cx.c[FMC_L6], cx.c[FMC_L5]);
    q = fma(r, q, -0.25);
    q = fma(r, q, cx.c[FMC_L3]);
    q = fma(r, q, -0.5);
    const double w = fma(kd, cx.c[FMC_LN2HI], logc);
    const double s = fma(kd, cx.c[FMC_LN2LO], r);
    L = w + fma(r2, q, s);
    // 1/z = invc / (1 + r) = invc (1 - r) (1 + r^2 + r^4 + r^6), |r|^8 < 6e-20
    double u = fma(r2, r2, r2);                   // r^2 + r^4
    u = fma(u, r2, r2);                           // r^2 + r^4 + r^6
    const double y0 = fma(-invc, r, invc);
    const double y = fma(y0, u, y0);
    rcp = fm_make(fm_hi(y) - ke, fm_lo(y));       // y * 2^-k
}

// N independent fast_log_rcp evaluations, written stage by stage so that the N dependency
// chains are interleaved in program order (DFMA latency on sm_100 is ~8 cycles at one issue
// per 2 cycles: four chains in flight per warp keep the FP64 pipe fed without relying on the
// scheduler to find the parallelism across separately inlined calls).  Unchecked: the caller
// guarantees 1e-300 <= t[i] <= 1e300.
// k ln2 + logc is ONE fma with ln2 rounded to double: z lies in [0.6875, 1.375), so k != 0
// implies |log t| >= 0.318 and the rounding error of ln2 (2.3e-17 |k|) stays below half an
// ulp of the result; the hi/lo split of the scalar form buys nothing here.  12 FP64
// instructions per element (scalar form: 14); k enters as (double)(k 2^20) times ln2 2^-20.
template <int N>
GLMMA_HD void fast_log_rcp_vec(const double* t, const MathCtx& cx, double* L, double* rcp) {
    int ke[N];
    double z[N], kd[N], invc[N], logc[N], r[N], r2[N], q[N], u[N];
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int i = 0; i < N; ++i) {
        const int hi = fm_hi(t[i]);
        const int tmp = hi - 0x3FE60000;
        const int idx = (tmp >> 13) & 127;
        ke[i] = tmp & 0xfff00000;
        kd[i] = (double)ke[i];                      // k * 2^20: the shift is folded into the constant below
        z[i] = fm_make(hi - ke[i], fm_lo(t[i]));
#ifdef __CUDA_ARCH__
        (void)idx;
        const unsigned off = (((unsigned)tmp >> 6) & 0x3F80u) | cx.lt_lane;       // 128 bytes per entry
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(invc[i]), "=d"(logc[i]) : "r"(cx.lt_base + off));
#else
        invc[i] = cx.logtab[FM_LOG_STRIDE * idx]; logc[i] = cx.logtab[FM_LOG_STRIDE * idx + 1];
#endif
    }
    const double c6 = cx.c[FMC_L6], c5 = cx.c[FMC_L5], c3 = cx.c[FMC_L3], ln2s = cx.c[FMC_LN2S];
#define FM_VEC for (int i = 0; i < N; ++i)
#ifdef __CUDACC__
#define FM_UNROLL _Pragma("unroll")
#else
#define FM_UNROLL
#endif
    FM_UNROLL FM_VEC r[i] = fma(z[i], invc[i], -1.0);
    FM_UNROLL FM_VEC r2[i] = r[i] * r[i];
    FM_UNROLL FM_VEC q[i] = fma(r[i], c6, c5);
    FM_UNROLL FM_VEC logc[i] = fma(kd[i], ln2s, logc[i]);        // w = k ln2 + logc
    FM_UNROLL FM_VEC q[i] = fma(r[i], q[i], -0.25);
    FM_UNROLL FM_VEC u[i] = fma(r2[i], r2[i], r2[i]);            // r^2 + r^4
    FM_UNROLL FM_VEC q[i] = fma(r[i], q[i], c3);
    FM_UNROLL FM_VEC q[i] = fma(r[i], q[i], -0.5);
    FM_UNROLL FM_VEC invc[i] = fma(-invc[i], r[i], invc[i]);     // y0 = invc (1 - r)
    FM_UNROLL FM_VEC q[i] = fma(r2[i], q[i], r[i]);
    // 1/z = y0 (1 + r^2 + r^4) up to r^6 <= 3.6e-15 relative: the reciprocal only feeds score
    // carriers (tolerance 1e-10), the logarithm keeps its full series
    FM_UNROLL FM_VEC invc[i] = fma(invc[i], u[i], invc[i]);      // y
    FM_UNROLL FM_VEC L[i] = logc[i] + q[i];
    FM_UNROLL FM_VEC rcp[i] = fm_make(fm_hi(invc[i]) - ke[i], fm_lo(invc[i]));
}

// N independent fast_exp evaluations, stage by stage (see fast_log_rcp_vec).  Unchecked:
// the caller guarantees |x[i]| <= 700.
template <int N>
GLMMA_HD void fast_exp_vec(const double* x, const MathCtx& cx, double* out) {
    int n[N];
    double kf[N], r[N], r2[N], p[N], tj[N];
    const double c64 = cx.c[FMC_E64LN2], mg = cx.c[FMC_EMAGIC], ehi = cx.c[FMC_EHI], elo = cx.c[FMC_ELO];
    const double e120 = cx.c[FMC_E120], e24 = cx.c[FMC_E24], e6 = cx.c[FMC_E6];
    FM_UNROLL FM_VEC kf[i] = fma(x[i], c64, mg);
    FM_UNROLL FM_VEC { n[i] = fm_lo(kf[i]); tj[i] = cx.exptab[FM_EXP_STRIDE * (n[i] & 63)]; }
    FM_UNROLL FM_VEC kf[i] = kf[i] - mg;
    FM_UNROLL FM_VEC r[i] = fma(kf[i], ehi, x[i]);
    FM_UNROLL FM_VEC r[i] = fma(kf[i], elo, r[i]);
    FM_UNROLL FM_VEC r2[i] = r[i] * r[i];
    FM_UNROLL FM_VEC p[i] = fma(r[i], e120, e24);
    FM_UNROLL FM_VEC p[i] = fma(r[i], p[i], e6);
    FM_UNROLL FM_VEC p[i] = fma(r[i], p[i], 0.5);
    FM_UNROLL FM_VEC p[i] = fma(r2[i], p[i], r[i]);
    FM_UNROLL FM_VEC p[i] = fma(tj[i], p[i], tj[i]);
    FM_UNROLL FM_VEC out[i] = fm_make(fm_hi(p[i]) + ((n[i] >> 6) << 20), fm_lo(p[i]));
#undef FM_VEC
#undef FM_UNROLL
}

// exp x = 2^(n/64) exp(r), n = rint(64 x / ln2), |r| <= ln2/128
GLMMA_HD double fast_exp(double x, const MathCtx& cx) {
    if (!(fabs(x) <= 700.0)) return fm_slow_exp(x);
    const double kf = fma(x, cx.c[FMC_E64LN2], cx.c[FMC_EMAGIC]);
    const int n = fm_lo(kf);
    const double nd = kf - cx.c[FMC_EMAGIC];
    double r = fma(nd, cx.c[FMC_EHI], x);
    r = fma(nd, cx.c[FMC_ELO], r);
    const double r2 = r * r;
    double p = fma(r, cx.c[FMC_E120], cx.c[FMC_E24]);
    p = fma(r, p, cx.c[FMC_E6]);
    p = fma(r, p, 0.5);
    p = fma(r2, p, r);
    const double tj = cx.exptab[FM_EXP_STRIDE * (n & 63)];
    const double v = fma(tj, p, tj);
    return fm_make(fm_hi(v) + ((n >> 6) << 20), fm_lo(v));
}


// ------------------------------------------------------------------------------------------
// Gaussian tail functions for censored.normal (R/Fit_Funs.R:1360-1435): log Phi(x) and the inverse
// Mills ratio phi(x) / Phi(x), without libm's erfc / erfcx / log / exp (about 600 instructions per
// censored point and, inlined eight times, an instruction-cache problem: profiles/README.md).
//
//   erfcx(u) = exp(u^2) erfc(u), u >= 0:   (1 + 2u) erfcx(u) = p26(t),  t = (u - 4) / (u + 4)
// p26: Chebyshev interpolant of degree 26 converted to the monomial basis (coefficients of
// magnitude <= 1.24; generated with mpmath at 60 digits, truncation error 2e-19, fp64 evaluation
// error <= 4 ulp -- tests/test_fastmath.py).
//   x <= 0:  Phi(x) = erfcx(-x/sqrt2) exp(-x^2/2) / 2  =>  log Phi = -x^2/2 - ln2 + log c,  phi/Phi = sqrt(2/pi) / c
//   x >  0:  Phi(x) = 1 - erfcx(x/sqrt2) exp(-x^2/2) / 2 = P  =>  log Phi = log P,  phi/Phi = exp(-x^2/2) / (sqrt(2 pi) P)
// one fast_log_rcp serves the logarithm and the reciprocal in both cases.
// ------------------------------------------------------------------------------------------
#define GLMMA_ERFCX_DEG 26
#ifdef __CUDA_ARCH__
static __constant__ double GLMMA_ERFCX_C[GLMMA_ERFCX_DEG + 1] = {
#else
static const double GLMMA_ERFCX_C[GLMMA_ERFCX_DEG + 1] = {
#endif
    1.23299511862555256e+00,
    -1.39621116840562498e-01,
    1.53796521026200415e-02,
    6.80970542546903562e-02,
    -1.01039066036325037e-01,
    9.37328349984395126e-02,
    -6.63303658116485007e-02,
    3.71675155359114107e-02,
    -1.61977340659525300e-02,
    5.03196999281092534e-03,
    -7.57773219535229294e-04,
    -1.99256776457919693e-04,
    1.50621307463917777e-04,
    -2.43989423090698060e-05,
    -1.12207534527254248e-05,
    5.70920121222635183e-06,
    2.90906296837142686e-07,
    -8.24912743496457224e-07,
    7.89343963476141997e-08,
    1.10090415607283036e-07,
    -2.19171286847767334e-08,
    -1.47051875523876298e-08,
    3.97990820993533376e-09,
    1.78874921890775359e-09,
    -5.47796315403793008e-10,
    -1.37212432109557996e-10,
    4.33574469688012796e-11};

// 1 / t, unchecked (1e-300 <= t <= 1e300): the reciprocal half of fast_log_rcp (6 FP64 instructions)
GLMMA_HD double fast_rcp(double t, const MathCtx& cx) {
    const int hi = fm_hi(t);
    const int tmp = hi - 0x3FE60000;
    const int i = (tmp >> 13) & 127;
    const int ke = tmp & 0xfff00000;
    const double z = fm_make(hi - ke, fm_lo(t));
    const double invc = cx.logtab[FM_LOG_STRIDE * i];
    const double r = fma(z, invc, -1.0);
    const double r2 = r * r;
    double u = fma(r2, r2, r2);
    u = fma(u, r2, r2);
    const double y0 = fma(-invc, r, invc);
    const double y = fma(y0, u, y0);
    return fm_make(fm_hi(y) - ke, fm_lo(y));
}

// erfcx(u) for 0 <= u <= 1e100
GLMMA_HD double fast_erfcx(double u, const MathCtx& cx) {
    const double d1 = u + 4.0, d2 = fma(2.0, u, 1.0);
    const double rp = fast_rcp(d1 * d2, cx);          // one reciprocal for both denominators
    const double t = (u - 4.0) * (rp * d2);            // (u - 4) / (u + 4)
    const double t2 = t * t;
    // even / odd halves: two Horner chains of 13 instead of one of 26
    double pe = GLMMA_ERFCX_C[26], po = GLMMA_ERFCX_C[25];
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 24; k >= 0; k -= 2) {
        pe = fma(pe, t2, GLMMA_ERFCX_C[k]);
        if (k >= 2) po = fma(po, t2, GLMMA_ERFCX_C[k - 1]);
    }
    return fma(t, po, pe) * (rp * d1);                 // / (1 + 2u)
}

// log Phi(x), phi(x) / Phi(x), phi(x) and Phi(x); branch-free, |x| <= 37 (beyond, exp takes its
// library path and the results stay correct: phi and, for x < 0, Phi underflow to 0)
GLMMA_HD void fast_gauss_tail(double x, const MathCtx& cx, double& logphi, double& mills, double& dens, double& cdf) {
    const double c = fast_erfcx(fabs(x) * 0.70710678118654752, cx);
    const double hx = -0.5 * x;
    const double h = hx * x;
    const double hl = fma(hx, x, -h);                  // exact rounding residual of -x^2/2
    double E = fast_exp(h, cx);
    E = fma(E, hl, E);                                 // exp(h + hl): keeps exp(-x^2/2) at ~3 ulp for large |x|
    const bool neg = x <= 0.0;
    const double q = 0.5 * E * c;                      // Phi(-|x|)
    const double P = 1.0 - q;                          // Phi(|x|)
    double L, r;
    fast_log_rcp<false>(neg ? c : P, cx, L, r);
    logphi = neg ? (h - 0.69314718055994531) + L : L;
    dens = 0.39894228040143268 * E;
    mills = neg ? 0.79788456080286536 * r : dens * r;
    cdf = neg ? q : P;
}
GLMMA_HD void fast_logphi_mills(double x, const MathCtx& cx, double& logphi, double& mills) {
    double dens, cdf;
    fast_gauss_tail(x, cx, logphi, mills, dens, cdf);
}


// two independent fast_log_rcp evaluations (full series for both results), interleaved; unchecked
GLMMA_HD void fast_log_rcp_pair(const double* t, const MathCtx& cx, double* L, double* rcp) {
    double L0, r0, L1, r1;
    fast_log_rcp<false>(t[0], cx, L0, r0);
    fast_log_rcp<false>(t[1], cx, L1, r1);
    L[0] = L0; rcp[0] = r0; L[1] = L1; rcp[1] = r1;
}

// ------------------------------------------------------------------------------------------
// lgamma(x) and digamma(x) together, 1e-290 <= x <= 1e25 (beta.fam / hurdle.beta.fam evaluate both at
// a = mu phi and b = (1 - mu) phi at every quadrature point; libm's lgamma + a digamma with a
// data-dependent shift loop cost ~1000 instructions per point there).
//   shift by 10:  P(x) = x (x+1) ... (x+9) = s (s+8) (s+14) (s+18) (s+20),  s = x (x + 9)
//   lgamma(x)  = lgamma(x + 10) - log P,   digamma(x) = digamma(x + 10) - P'(x) / P(x)
//   Stirling series at z = x + 10 >= 10 (7 terms: truncation < 1e-16 absolute)
// Two fast_log_rcp (z and P) supply both logarithms and both reciprocals.  Error: lgamma 12 ulp or
// 1.5e-14 absolute (a difference of two terms of size 13..25), digamma 8 ulp or 4e-15 absolute
// (tests/test_fastmath.py).
// ------------------------------------------------------------------------------------------
GLMMA_HD void fast_lgamma_digamma(double x, const MathCtx& cx, double& lg, double& psi) {
    const double s = x * (x + 9.0);
    const double f1 = s + 8.0, f2 = s + 14.0, f3 = s + 18.0, f4 = s + 20.0;
    const double t34 = f3 * f4, s34 = f3 + f4;
    const double u2 = fma(f2, s34, t34);          // d/ds of f2 f3 f4
    const double p234 = f2 * t34;
    const double u1 = fma(f1, u2, p234);          // d/ds of f1 f2 f3 f4
    const double p1234 = f1 * p234;
    const double u0 = fma(s, u1, p1234);          // dP/ds
    const double P = s * p1234;
    const double dP = u0 * fma(2.0, x, 9.0);      // P'(x)
    const double z = x + 10.0;
    double tt[2] = {z, P}, LL[2], rr[2];
    fast_log_rcp_pair(tt, cx, LL, rr);
    const double Lz = LL[0], rz = rr[0], w = rz * rz;
    double a = fma(w, 6.41025641025641025641e-03, -1.91752691752691752692e-03);    // 1/156, -691/360360
    a = fma(w, a, 8.41750841750841750842e-04);     // 1/1188
    a = fma(w, a, -5.95238095238095238095e-04);    // -1/1680
    a = fma(w, a, 7.93650793650793650794e-04);     // 1/1260
    a = fma(w, a, -2.77777777777777777778e-03);    // -1/360
    a = fma(w, a, 8.33333333333333333333e-02);     // 1/12
    const double lgz = fma(z - 0.5, Lz, fma(rz, a, 0.91893853320467274178 - z));
    double b = fma(w, -8.33333333333333333333e-02, 2.10927960927960927961e-02);      // -1/12, 691/32760
    b = fma(w, b, -7.57575757575757575758e-03);    // -1/132
    b = fma(w, b, 4.16666666666666666667e-03);     // 1/240
    b = fma(w, b, -3.96825396825396825397e-03);    // -1/252
    b = fma(w, b, 8.33333333333333333333e-03);     // 1/120
    b = fma(w, b, -8.33333333333333333333e-02);    // -1/12
    const double psz = fma(w, b, fma(-0.5, rz, Lz));
    lg = lgz - LL[1];
    psi = fma(-dP, rr[1], psz);
}

// 1 - exp(-x) for 0 <= x: series below 2^-10 (truncation x^5 / 720 relative), 1 - fast_exp(-x) above
// (2.3e-13 relative at the switch, improving with x).  Replaces -expm1(-x) in the truncated-at-zero
// count densities and the complementary log-log link.
GLMMA_HD double fast_one_minus_expneg(double x, const MathCtx& cx) {
    if (x < 9.765625e-4)
        return x * fma(x, fma(x, fma(x, fma(x, 8.33333333333333333333e-03, -4.16666666666666666667e-02),
                                     1.66666666666666666667e-01), -0.5), 1.0);
    return 1.0 - fast_exp(-x, cx);
}

// log(1 + exp(x)) and plogis(x) from exp(x) = ex (any finite positive value, or 0/inf)
// CHECK = false: the caller guarantees 1e-300 <= ex <= 1e300
template <bool CHECK = true>
GLMMA_HD void fast_l1p_sigmoid(double ex, const MathCtx& cx, double& l1p, double& sig) {
    if (ex < 4503599627370496.0) {           // 2^52: below, 1 + ex still resolves ex
        double L, r;
        fast_log_rcp<CHECK>(1.0 + ex, cx, L, r);
        // for tiny ex, log(1 + ex) rounds like ex - ex^2/2: the table result is within 2e-16 absolute
        l1p = L;
        sig = ex * r;
    } else {
        l1p = fm_slow_log(ex);
        sig = 1.0;
    }
}
