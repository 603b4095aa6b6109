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
// instruction immediates.  The kernels load the constants from shared memory once, so
// they live in registers; as literals ptxas re-materialises each of them with two
// moves at every use (ncu: 24 of 134 instructions per 32 points were such moves).
struct MathCtx {
    const double* logtab;   // 128 x (invc, logc)
    const double* exptab;   // 64 x 2^(j/64)
    double l6, l5, l3, ln2hi, ln2lo;                          // -1/6, 1/5, 1/3, ln2 split
    double e64ln2, emagic, ehi, elo, e120, e24, e6;           // exp: 64/ln2, 1.5*2^52, -ln2/64 split, 1/5!, 1/4!, 1/3!
};
#define GLMMA_FM_NCONST 12
// c must hold the GLMMA_FM_NCONST constants in this order (fastmath_constants())
GLMMA_HD MathCtx make_mathctx(const double* logtab, const double* exptab, const double* c) {
    MathCtx m;
    m.logtab = logtab; m.exptab = exptab;
    m.l6 = c[0]; m.l5 = c[1]; m.l3 = c[2]; m.ln2hi = c[3]; m.ln2lo = c[4];
    m.e64ln2 = c[5]; m.emagic = c[6]; m.ehi = c[7]; m.elo = c[8]; m.e120 = c[9]; m.e24 = c[10]; m.e6 = c[11];
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
    const double2 tc = *reinterpret_cast<const double2*>(cx.logtab + 2 * i);
    const double invc = tc.x, logc = tc.y;
#else
    const double invc = cx.logtab[2 * i], logc = cx.logtab[2 * i + 1];
#endif
    const double kd = (double)k;
    const double r = fma(z, invc, -1.0);
    const double r2 = r * r;
    double q = fma(r, cx.l6, cx.l5);
    q = fma(r, q, -0.25);
    q = fma(r, q, cx.l3);
    q = fma(r, q, -0.5);
    const double w = fma(kd, cx.ln2hi, logc);
    const double s = fma(kd, cx.ln2lo, r);
    L = w + fma(r2, q, s);
    // 1/z = invc / (1 + r) = invc (1 - r) (1 + r^2 + r^4 + r^6), |r|^8 < 6e-20
    double u = fma(r2, r2, r2);                   // r^2 + r^4
    u = fma(u, r2, r2);                           // r^2 + r^4 + r^6
    const double y0 = fma(-invc, r, invc);
    const double y = fma(y0, u, y0);
    rcp = fm_make(fm_hi(y) - ke, fm_lo(y));       // y * 2^-k
}

// exp x = 2^(n/64) exp(r), n = rint(64 x / ln2), |r| <= ln2/128
GLMMA_HD double fast_exp(double x, const MathCtx& cx) {
    if (!(fabs(x) <= 700.0)) return fm_slow_exp(x);
    const double kf = fma(x, cx.e64ln2, cx.emagic);
    const int n = fm_lo(kf);
    const double nd = kf - cx.emagic;
    double r = fma(nd, cx.ehi, x);
    r = fma(nd, cx.elo, r);
    const double r2 = r * r;
    double p = fma(r, cx.e120, cx.e24);
    p = fma(r, p, cx.e6);
    p = fma(r, p, 0.5);
    p = fma(r2, p, r);
    const double tj = cx.exptab[n & 63];
    const double v = fma(tj, p, tj);
    return fm_make(fm_hi(v) + ((n >> 6) << 20), fm_lo(v));
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
