"""Restatement of the pieces of R's RNG needed to regenerate the vignette data
sets that carry the reference's published known-answer values (TEST
INFRASTRUCTURE ONLY; see oracle/__init__.py).

R's default generators (R >= 3.6): Mersenne-Twister + Inversion.  R's own
sources are not under /root/reference; the recipe is SURVEY.md appendix E and
is verified in tests/test_oracle_kat.py against the documented outputs of
``set.seed(1234); runif(3)`` and ``set.seed(42); rnorm(3)``.
Call sites in the reference: vignettes/GLMMadaptive.Rmd:82-107,
vignettes/ZeroInflated_and_TwoPart_Models.Rmd:143-178,313-351.
"""
import numpy as np
from scipy.special import ndtri

_I2_32M1 = 2.328306437080797e-10  # 1/(2^32 - 1)


class RRng:
    def __init__(self, seed):
        # RNG.c: Randomize(): initial scrambling of the seed, then FixupSeeds
        s = np.uint32(seed & 0xFFFFFFFF)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(69069) * s + np.uint32(1)
            iseed = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = np.uint32(69069) * s + np.uint32(1)
                iseed[j] = s
        # i_seed[0] is dummy[0] = mti, forced to N=624 by FixupSeeds
        self._bg = np.random.MT19937()
        st = self._bg.state
        st["state"]["key"] = iseed[1:].copy()
        st["state"]["pos"] = 624
        self._bg.state = st

    def unif_rand(self, n=None):
        raw = self._bg.random_raw(1 if n is None else n).astype(np.float64)
        u = raw * 2.3283064365386963e-10
        # fixup(): keep strictly inside (0, 1)
        u = np.where(u <= 0.0, 0.5 * _I2_32M1, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * _I2_32M1, u)
        return u[0] if n is None else u

    def runif(self, n, a=0.0, b=1.0):
        a = np.broadcast_to(np.asarray(a, float), (n,))
        b = np.broadcast_to(np.asarray(b, float), (n,))
        return a + (b - a) * self.unif_rand(n)

    def norm_rand(self, n):
        # INVERSION: u = unif_rand(); u = (int)(BIG*u) + unif_rand(); qnorm5(u/BIG)
        BIG = 134217728.0
        u = self.unif_rand(2 * n)
        u1, u2 = u[0::2], u[1::2]
        v = np.floor(BIG * u1) + u2
        return ndtri(v / BIG)

    def rnorm(self, n, mean=0.0, sd=1.0):
        mean = np.broadcast_to(np.asarray(mean, float), (n,))
        sd = np.broadcast_to(np.asarray(sd, float), (n,))
        return mean + sd * self.norm_rand(n)

    def rbinom1(self, n, prob):
        """rbinom(n, size = 1, prob): with n*min(p,1-p) < 30 R uses the inversion
        branch, which for size 1 consumes exactly one uniform per draw."""
        prob = np.broadcast_to(np.asarray(prob, float), (n,))
        u = self.unif_rand(n)
        p = np.minimum(prob, 1.0 - prob)
        q = 1.0 - p
        # inversion: ix = 0; f = q^1 = q; if u < f -> 0 else 1 (for size 1)
        ix = (u >= q).astype(float)
        # psave > 0.5 -> ix = n - ix
        return np.where(prob > 0.5, 1.0 - ix, ix)


# ----------------------------------------------------------------------------------------------
# Scalar generators of R's nmath (R >= 3.6 defaults), restated from their published algorithms.
# R's sources are not under /root/reference; each routine names the paper R implements.  They
# are pinned in tests/test_oracle_kat.py: with these generators the vignette data sets are
# regenerated and the reference's PUBLISHED log-likelihoods are reproduced at the published
# estimates (poisson gm1, zi.poisson fm1/fm2, zi.nb gm1, negative binomial fm1, beta gm) --
# one wrong branch or constant anywhere below changes the data and breaks those numbers.
# Call sites in the reference: vignettes/GLMMadaptive.Rmd:225 (rpois),
# vignettes/ZeroInflated_and_TwoPart_Models.Rmd:76 (rnbinom), vignettes/Custom_Models.Rmd:74
# (rnbinom, size < 1), :287 (rbeta).
# ----------------------------------------------------------------------------------------------
import math

_M_1_SQRT_2PI = 0.398942280401432677939946059934


class RScalarRng(RRng):
    """RRng plus the per-draw state R's C routines keep in static variables."""

    def __init__(self, seed):
        super().__init__(seed)
        # exp_rand(): q[k-1] = sum_{j<=k} log(2)^j / j!
        q, term, acc = [], 1.0, 0.0
        for k in range(1, 17):
            term *= math.log(2.0) / k
            acc += term
            q.append(acc)
        q[15] = 1.0
        self._q = q
        self._pois = dict(muprev=0.0, muprev2=0.0, l=0, m=0, pp=[0.0] * 36)
        self._gam = dict(aa=0.0, aaa=0.0)
        self._beta = dict(olda=-1.0, oldb=-1.0)

    def u(self):
        return float(self.unif_rand())

    def norm1(self):
        BIG = 134217728.0
        u1 = self.u()
        u1 = float(int(BIG * u1)) + self.u()
        return float(ndtri(u1 / BIG))

    def exp_rand(self):
        """Ahrens & Dieter (1972) algorithm SA, as in nmath/sexp.c."""
        q = self._q
        a = 0.0
        u = self.u()
        while u <= 0.0 or u >= 1.0:
            u = self.u()
        while True:
            u += u
            if u > 1.0:
                break
            a += q[0]
        u -= 1.0
        if u <= q[0]:
            return a + u
        i = 0
        ustar = self.u()
        umin = ustar
        while True:
            ustar = self.u()
            if umin > ustar:
                umin = ustar
            i += 1
            if not (u > q[i]):
                break
        return a + umin * q[0]

    # ---- rpois: Ahrens & Dieter (1982) algorithm PD, nmath/rpois.c --------------------------
    def rpois1(self, mu):
        a0, a1, a2, a3 = -0.5, 0.3333333, -0.2500068, 0.2000118
        a4, a5, a6, a7 = -0.1661269, 0.1421878, -0.1384794, 0.1250060
        one_7, one_12, one_24 = 0.1428571428571428571, 0.0833333333333333333, 0.0416666666666666667
        fact = [1., 1., 2., 6., 24., 120., 720., 5040., 40320., 362880.]
        S = self._pois
        if not math.isfinite(mu) or mu < 0:
            return float("nan")
        if mu <= 0.0:
            return 0.0
        big_mu = mu >= 10.0
        new_big_mu = False
        if not (big_mu and mu == S["muprev"]):
            if big_mu:
                new_big_mu = True
                S["muprev"] = mu
                S["s"] = math.sqrt(mu)
                S["d"] = 6.0 * mu * mu
                S["big_l"] = math.floor(mu - 1.1484)
            else:
                if mu != S["muprev"]:
                    S["muprev"] = mu
                    S["m"] = max(1, int(mu))
                    S["l"] = 0
                    S["q"] = S["p0"] = S["p"] = math.exp(-mu)
                pp = S["pp"]
                while True:
                    u = self.u()
                    if u <= S["p0"]:
                        return 0.0
                    if S["l"] > 0:
                        j = min(S["l"], S["m"]) if u > 0.458 else 1
                        for k in range(j, S["l"] + 1):
                            if u <= pp[k]:
                                return float(k)
                        if S["l"] == 35:
                            continue
                    S["l"] += 1
                    for k in range(S["l"], 36):
                        S["p"] *= mu / k
                        S["q"] += S["p"]
                        pp[k] = S["q"]
                        if u <= S["q"]:
                            S["l"] = k
                            return float(k)
                    S["l"] = 35
        s, d, big_l = S["s"], S["d"], S["big_l"]
        # Step N
        g = mu + s * self.norm1()
        pois = -1.0
        u = 0.0
        fk = difmuk = 0.0
        if g >= 0.0:
            pois = math.floor(g)
            if pois >= big_l:
                return pois
            fk = pois
            difmuk = mu - fk
            u = self.u()
            if d * u >= difmuk * difmuk * difmuk:
                return pois
        # Step P
        if new_big_mu or mu != S["muprev2"]:
            S["muprev2"] = mu
            S["omega"] = _M_1_SQRT_2PI / s
            b1 = one_24 / mu
            b2 = 0.3 * b1 * b1
            c3 = one_7 * b1 * b2
            S["c3"] = c3
            S["c2"] = b2 - 15.0 * c3
            S["c1"] = b1 - 6.0 * b2 + 45.0 * c3
            S["c0"] = 1.0 - b1 + 3.0 * b2 - 15.0 * c3
            S["c"] = 0.1069 / mu
        omega, c3, c2, c1, c0, c = S["omega"], S["c3"], S["c2"], S["c1"], S["c0"], S["c"]

        def step_F(pois, fk, difmuk):
            if pois < 10:
                px = -mu
                py = mu ** pois / fact[int(pois)]
            else:
                dl = one_12 / fk
                dl = dl - 4.8 * dl * dl * dl
                v = difmuk / fk
                if abs(v) <= 0.25:
                    px = fk * v * v * (((((((a7 * v + a6) * v + a5) * v + a4) * v + a3) * v + a2)
                                        * v + a1) * v + a0) - dl
                else:
                    px = fk * math.log(1.0 + v) - difmuk - dl
                py = _M_1_SQRT_2PI / math.sqrt(fk)
            x = (0.5 - difmuk) / s
            x *= x
            fx = -0.5 * x
            fy = omega * (((c3 * x + c2) * x + c1) * x + c0)
            return px, py, fx, fy

        if g >= 0.0:
            px, py, fx, fy = step_F(pois, fk, difmuk)
            if fy - u * fy <= py * math.exp(px - fx):      # Step Q
                return pois
        while True:
            E = self.exp_rand()
            u = 2.0 * self.u() - 1.0
            t = 1.8 + math.copysign(E, u)
            if t > -0.6744:
                pois = math.floor(mu + s * t)
                fk = pois
                difmuk = mu - fk
                px, py, fx, fy = step_F(pois, fk, difmuk)
                if c * abs(u) <= py * math.exp(px + E) - fy * math.exp(fx + E):   # Step H
                    return pois

    # ---- rgamma: Ahrens & Dieter GD (1982) for a >= 1, GS (1974) for a < 1; nmath/rgamma.c ----
    def rgamma1(self, a, scale):
        sqrt32, exp_m1 = 5.656854, 0.36787944117144233
        q1, q2, q3, q4 = 0.04166669, 0.02083148, 0.00801191, 0.00144121
        q5, q6, q7 = -7.388e-5, 2.4511e-4, 2.424e-4
        a1, a2, a3, a4 = 0.3333333, -0.250003, 0.2000062, -0.1662921
        a5, a6, a7 = 0.1423657, -0.1367177, 0.1233795
        S = self._gam
        if math.isnan(a) or math.isnan(scale):
            return float("nan")
        if a <= 0.0 or scale <= 0.0:
            if scale == 0.0 or a == 0.0:
                return 0.0
            return float("nan")
        if a < 1:
            e = 1.0 + exp_m1 * a
            while True:
                p = e * self.u()
                if p >= 1.0:
                    x = -math.log((e - p) / a)
                    if self.exp_rand() >= (1.0 - a) * math.log(x):
                        break
                else:
                    x = math.exp(math.log(p) / a)
                    if self.exp_rand() >= x:
                        break
            return scale * x
        if a != S["aa"]:
            S["aa"] = a
            S["s2"] = a - 0.5
            S["s"] = math.sqrt(S["s2"])
            S["d"] = sqrt32 - S["s"] * 12
        s, s2, d = S["s"], S["s2"], S["d"]
        t = self.norm1()
        x = s + 0.5 * t
        ret = x * x
        if t >= 0:
            return scale * ret
        u = self.u()
        if d * u <= t * t * t:
            return scale * ret
        if a != S["aaa"]:
            S["aaa"] = a
            r = 1 / a
            S["q0"] = ((((((q7 * r + q6) * r + q5) * r + q4) * r + q3) * r + q2) * r + q1) * r
            if a <= 3.686:
                S["b"] = 0.463 + s + 0.178 * s2
                S["si"] = 1.235
                S["c"] = 0.195 / s - 0.079 + 0.16 * s
            elif a <= 13.022:
                S["b"] = 1.654 + 0.0076 * s2
                S["si"] = 1.68 / s + 0.275
                S["c"] = 0.062 / s + 0.024
            else:
                S["b"] = 1.77
                S["si"] = 0.75
                S["c"] = 0.1515 / s
        q0, b, si, c = S["q0"], S["b"], S["si"], S["c"]

        def quot(t):
            v = t / (s + s)
            if abs(v) <= 0.25:
                return q0 + 0.5 * t * t * ((((((a7 * v + a6) * v + a5) * v + a4) * v + a3) * v
                                             + a2) * v + a1) * v
            return q0 - s * t + 0.25 * t * t + (s2 + s2) * math.log(1.0 + v)

        if x > 0.0:
            if math.log(1.0 - u) <= quot(t):
                return scale * ret
        while True:
            e = self.exp_rand()
            u = self.u()
            u = u + u - 1.0
            t = b - si * e if u < 0.0 else b + si * e
            if t >= -0.71874483771719:
                q = quot(t)
                if q > 0.0:
                    w = math.expm1(q)
                    if c * abs(u) <= w * math.exp(e - 0.5 * t * t):
                        break
        x = s + 0.5 * t
        return scale * x * x

    def rnbinom_mu1(self, size, mu):
        """nmath/rnbinom.c rnbinom_mu: a gamma mixture of Poissons."""
        return 0.0 if mu == 0 else self.rpois1(self.rgamma1(size, mu / size))

    # ---- rbeta: Cheng (1978) algorithms BB (min(a,b) > 1) and BC; nmath/rbeta.c ---------------
    def rbeta1(self, aa, bb):
        expmax = 1024 * math.log(2.0)
        DBL_MAX = 1.7976931348623157e308
        S = self._beta
        qsame = (S["olda"] == aa) and (S["oldb"] == bb)
        if not qsame:
            S["olda"], S["oldb"] = aa, bb
        a, b = min(aa, bb), max(aa, bb)
        alpha = a + b

        def v_w(u1, AA):
            v = S["beta"] * math.log(u1 / (1.0 - u1))
            if v <= expmax:
                try:
                    w = AA * math.exp(v)
                except OverflowError:
                    w = DBL_MAX
                if not math.isfinite(w):
                    w = DBL_MAX
            else:
                w = DBL_MAX
            return v, w

        if a <= 1.0:        # BC
            if not qsame:
                S["beta"] = 1.0 / a
                S["delta"] = 1.0 + b - a
                S["k1"] = S["delta"] * (0.0138889 + 0.0416667 * a) / (b * S["beta"] - 0.777778)
                S["k2"] = 0.25 + (0.5 + 0.25 / S["delta"]) * a
            while True:
                u1 = self.u()
                u2 = self.u()
                if u1 < 0.5:
                    y = u1 * u2
                    z = u1 * y
                    if 0.25 * u2 + z - y >= S["k1"]:
                        continue
                else:
                    z = u1 * u1 * u2
                    if z <= 0.25:
                        v, w = v_w(u1, b)
                        break
                    if z >= S["k2"]:
                        continue
                v, w = v_w(u1, b)
                if alpha * (math.log(alpha / (a + w)) + v) - 1.3862944 >= math.log(z):
                    break
            return a / (a + w) if aa == a else w / (a + w)
        # BB
        if not qsame:
            S["beta"] = math.sqrt((alpha - 2.0) / (2.0 * a * b - alpha))
            S["gamma"] = a + 1.0 / S["beta"]
        while True:
            u1 = self.u()
            u2 = self.u()
            v, w = v_w(u1, a)
            z = u1 * u1 * u2
            r = S["gamma"] * v - 1.3862944
            s = a + r - w
            if s + 2.609438 >= 5.0 * z:
                break
            t = math.log(z)
            if s > t:
                break
            if not (r + alpha * math.log(alpha / (b + w)) < t):
                break
        return b / (b + w) if aa != a else w / (b + w)

    # ---- vectorised front ends with R's recycling (one draw per element, in index order) -----
    def rpois(self, n, mu):
        mu = np.broadcast_to(np.asarray(mu, float), (n,))
        return np.array([self.rpois1(float(m)) for m in mu])

    def rnbinom_mu(self, n, size, mu):
        mu = np.broadcast_to(np.asarray(mu, float), (n,))
        size = np.broadcast_to(np.asarray(size, float), (n,))
        return np.array([self.rnbinom_mu1(float(s), float(m)) for s, m in zip(size, mu)])

    def rbeta(self, n, shape1, shape2):
        a = np.broadcast_to(np.asarray(shape1, float), (n,))
        b = np.broadcast_to(np.asarray(shape2, float), (n,))
        return np.array([self.rbeta1(float(x), float(y)) for x, y in zip(a, b)])
