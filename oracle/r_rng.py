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
