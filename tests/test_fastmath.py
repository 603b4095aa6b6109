"""Accuracy of the engine's fp64 special functions (glmmadaptive_b200/csrc/fastmath.cuh),
compiled for the host with g++ and compared with mpmath at 40 digits.  The same source
is what the kernels inline; fma() is correctly rounded on both sides."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest
from mpmath import mp, mpf, log as mplog, exp as mpexp

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm():
    out = os.path.join(tempfile.mkdtemp(prefix="glmma_fm_"), "libfm.so")
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC",
                    os.path.join(HERE, "fastmath_host.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    dp = C.POINTER(C.c_double)
    lib.hm_log_rcp.argtypes = [dp, C.c_int, dp, dp]
    lib.hm_exp.argtypes = [dp, C.c_int, dp]
    lib.hm_exp_vec3.argtypes = [dp, C.c_int, dp]
    lib.hm_log_rcp_vec4.argtypes = [dp, C.c_int, dp, dp]
    lib.hm_l1p_sigmoid.argtypes = [dp, C.c_int, dp, dp]
    lib.hm_logphi_mills.argtypes = [dp, C.c_int, dp, dp]
    lib.hm_erfcx.argtypes = [dp, C.c_int, dp]
    lib.hm_one_minus_expneg.argtypes = [dp, C.c_int, dp]
    lib.hm_lgamma_digamma.argtypes = [dp, C.c_int, dp, dp]
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def ulp(x):
    return np.spacing(np.abs(x))


def test_log_rcp(hm):
    mp.dps = 40
    rng = np.random.default_rng(0)
    t = np.concatenate([np.exp(rng.uniform(-690, 690, 20000)), rng.uniform(0.5, 2.0, 20000),
                        1.0 + rng.uniform(-1e-3, 1e-3, 5000), [1.0, 2.0, 0.5, 1e-300, 1e300, 0.6875, 1.375],
                        np.exp(rng.uniform(-5, 8, 20000))])
    L = np.empty_like(t); r = np.empty_like(t)
    hm.hm_log_rcp(_p(t), len(t), _p(L), _p(r))
    Lref = np.array([float(mplog(mpf(float(v)))) for v in t])
    rref = np.array([float(1 / mpf(float(v))) for v in t])
    eL = np.abs(L - Lref)
    assert np.all((eL <= 2 * ulp(Lref)) | (eL <= 2.3e-16)), (np.max(eL / ulp(Lref)), np.max(eL))
    assert np.max(np.abs(r - rref) / ulp(rref)) <= 2.0
    # the stage-wise 4-vector form used by the register-variant kernel is the same arithmetic
    n4 = len(t) // 4 * 4
    L4 = np.empty(n4); r4 = np.empty(n4)
    hm.hm_log_rcp_vec4(_p(t), n4, _p(L4), _p(r4))
    assert np.all(np.abs(L4 - Lref[:n4]) <= np.maximum(2 * ulp(Lref[:n4]), 2.3e-16))
    # its reciprocal stops at r^4 (score carriers only): 3.6e-15 relative
    assert np.max(np.abs(r4 - rref[:n4]) / np.abs(rref[:n4])) <= 4e-15
    # out-of-range arguments take the library path
    bad = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310])
    L = np.empty_like(bad); r = np.empty_like(bad)
    hm.hm_log_rcp(_p(bad), len(bad), _p(L), _p(r))
    with np.errstate(all="ignore"):
        np.testing.assert_array_equal(L, np.log(bad))
        np.testing.assert_array_equal(r, 1.0 / bad)


def test_exp(hm):
    mp.dps = 40
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-700, 700, 30000), rng.uniform(-40, 40, 30000),
                        rng.uniform(-1e-3, 1e-3, 3000), [0.0, -700.0, 700.0, 1e-320, -0.0]])
    out = np.empty_like(x)
    hm.hm_exp(_p(x), len(x), _p(out))
    ref = np.array([float(mpexp(mpf(float(v)))) for v in x])
    assert np.max(np.abs(out - ref) / ulp(ref)) <= 2.0
    # the stage-wise 3-vector form (table building of the register-variant kernel): same arithmetic
    n3 = len(x) // 3 * 3
    o3 = np.empty(n3)
    hm.hm_exp_vec3(_p(x), n3, _p(o3))
    np.testing.assert_array_equal(o3, out[:n3])
    far = np.array([-800.0, 720.0, np.nan, -np.inf, np.inf, -745.0])
    out = np.empty_like(far)
    hm.hm_exp(_p(far), len(far), _p(out))
    with np.errstate(all="ignore"):
        np.testing.assert_array_equal(out, np.exp(far))


def test_l1p_sigmoid(hm):
    mp.dps = 40
    rng = np.random.default_rng(2)
    ez = rng.uniform(-60, 60, 20000)
    ex = np.exp(ez)
    l1p = np.empty_like(ex); sig = np.empty_like(ex)
    hm.hm_l1p_sigmoid(_p(ex), len(ex), _p(l1p), _p(sig))
    ref_l = np.array([float(mplog(1 + mpf(float(v)))) for v in ex])
    ref_s = np.array([float(mpf(float(v)) / (1 + mpf(float(v)))) for v in ex])
    assert np.max(np.abs(l1p - ref_l)) <= 4e-16 + 2 * np.max(ulp(ref_l))
    assert np.all(np.abs(l1p - ref_l) <= np.maximum(3 * ulp(ref_l), 2.3e-16))
    assert np.all(np.abs(sig - ref_s) <= np.maximum(3 * ulp(ref_s), 1e-300))


def test_erfcx_logphi_mills(hm):
    """Gaussian tail functions of the censored-normal fast path against mpmath."""
    from mpmath import erfc, ncdf, npdf, sqrt as mpsqrt
    mp.dps = 50
    rng = np.random.default_rng(3)
    u = np.concatenate([rng.uniform(0, 8, 6000), rng.uniform(0, 0.05, 500), np.exp(rng.uniform(0, 6, 500)), [0.0, 4.0, 1e3]])
    out = np.empty_like(u)
    hm.hm_erfcx(_p(u), len(u), _p(out))

    def erfcx_ref(v):
        v = mpf(float(v))
        return mpexp(v * v) * erfc(v) if v < 20 else sum((-1) ** k * mp.rf(mpf(1) / 2, k) / v ** (2 * k) for k in range(30)) / (v * mpsqrt(mp.pi))
    ref = np.array([float(erfcx_ref(v)) for v in u])
    assert np.max(np.abs(out - ref) / ulp(ref)) <= 6.0
    x = np.concatenate([rng.uniform(-8, 8, 8000), rng.uniform(-0.01, 0.01, 300), [0.0, -30.0, 8.0, -4.5, 4.5]])
    lp = np.empty_like(x); ml = np.empty_like(x)
    hm.hm_logphi_mills(_p(x), len(x), _p(lp), _p(ml))
    lref = np.array([float(mplog(ncdf(mpf(float(v))))) for v in x])
    mref = np.array([float(npdf(mpf(float(v))) / ncdf(mpf(float(v)))) for v in x])
    # log Phi: 8 ulp or 3e-16 absolute (x > 0: log of a number next to 1); Mills ratio: 8 ulp
    assert np.all(np.abs(lp - lref) <= np.maximum(8 * ulp(lref), 3e-16)), np.max(np.abs(lp - lref))
    assert np.max(np.abs(ml - mref) / np.maximum(ulp(mref), 1e-300)) <= 8.0


def test_lgamma_digamma(hm):
    """Joint lgamma / digamma of the beta families' fast path against mpmath."""
    from mpmath import loggamma, digamma
    mp.dps = 50
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(0, 12, 6000), np.exp(rng.uniform(-25, 10, 4000)), rng.uniform(0.9, 2.1, 1500),
                        [1.0, 2.0, 1.4616321449683623, 1e-12, 0.5, 10.0, 1e4, 1e8]])
    lg = np.empty_like(x); ps = np.empty_like(x)
    hm.hm_lgamma_digamma(_p(x), len(x), _p(lg), _p(ps))
    lref = np.array([float(loggamma(mpf(float(v)))) for v in x])
    pref = np.array([float(digamma(mpf(float(v)))) for v in x])
    # lgamma(x + 10) - log P(x): the two terms are of size 13..25 for x < 12, so the absolute error is
    # a few ulps OF THOSE (1.5e-14), whatever the size of the result; 12 ulp for larger arguments
    assert np.all(np.abs(lg - lref) <= np.maximum(12 * ulp(lref), 1.5e-14)), np.max(np.abs(lg - lref))
    assert np.all(np.abs(ps - pref) <= np.maximum(8 * ulp(pref), 4e-15)), np.max(np.abs(ps - pref))


def test_one_minus_expneg(hm):
    """1 - exp(-x) of the zero-truncated count densities and the cloglog link: relative 3e-13."""
    from mpmath import expm1
    mp.dps = 40
    rng = np.random.default_rng(5)
    x = np.concatenate([np.exp(rng.uniform(-36, 6, 20000)), rng.uniform(5e-4, 2e-3, 2000), [9.765625e-4, 0.0, 700.0, 1e-300]])
    out = np.empty_like(x)
    hm.hm_one_minus_expneg(_p(x), len(x), _p(out))
    ref = np.array([float(-expm1(-mpf(float(v)))) for v in x])
    nz = ref > 0
    assert np.max(np.abs(out[nz] - ref[nz]) / ref[nz]) <= 3e-13
    assert out[~nz].tolist() == [0.0] * int((~nz).sum())
