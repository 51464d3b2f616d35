"""TEST INFRASTRUCTURE ONLY (see oracle/nmma_oracle.py header): NumPy restatement of the counter-based
generator behind ``nmma_b200_prior_sample`` -- Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11; Random123
``philox.h`` constants; the algorithm cuRAND and torch use).  Pinned by the Random123 known-answer vectors
(``kat_vectors``: counter/key all-zero, all-ones and the digits of pi) in ``tests/test_prior_device.py``.

Not part of the reference (bilby draws with NumPy's global generator, ``bilby/core/prior/base.py: sample``);
what must match the reference is the *transform* of a unit-cube point, restated in ``rescale_columns`` below
from bilby.core.prior.analytical (third-party, absent offline) in the operation order of its Python source.
"""
import numpy as np
from scipy.special import erf, erfinv

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over counter arrays (uint64 holding 32-bit words); returns four uint64 arrays of 32-bit words."""
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u01_53(a, b):
    return ((a >> np.uint64(5)) * np.uint64(1 << 26) + (b >> np.uint64(6))).astype(np.float64) / 9007199254740992.0


def unit_cube(seed, first_index, n, P):
    """unit[n, P]: point i, columns (2j, 2j+1) from Philox(counter=(idx_lo, idx_hi, j, 0), key=(seed_lo, seed_hi))."""
    idx = np.arange(first_index, first_index + n, dtype=np.uint64)
    out = np.empty((n, P))
    for j in range((P + 1) // 2):
        r = philox4x32_10(idx & MASK, idx >> np.uint64(32), np.full(n, j, np.uint64), np.zeros(n, np.uint64),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        out[:, 2 * j] = u01_53(r[0], r[1])
        if 2 * j + 1 < P:
            out[:, 2 * j + 1] = u01_53(r[2], r[3])
    return out


def rescale_column(kind, par, u, table=None):
    """bilby.core.prior.analytical.<Prior>.rescale, expression for expression."""
    u = np.asarray(u, float)
    if kind == "Uniform":
        mn, mx = par[:2]
        return mn + u * (mx - mn)
    if kind == "DeltaFunction":
        return par[0] * u ** 0
    if kind == "Sine":
        mn, mx = par[:2]
        norm = 1 / (np.cos(mn) - np.cos(mx))
        return np.arccos(np.cos(mn) - u / norm)
    if kind == "Cosine":
        mn, mx = par[:2]
        norm = 1 / (np.sin(mx) - np.sin(mn))
        return np.arcsin(u / norm + np.sin(mn))
    if kind == "Gaussian":
        mu, sigma = par[:2]
        return mu + erfinv(2 * u - 1) * 2 ** 0.5 * sigma
    if kind == "TruncatedGaussian":
        mu, sigma, mn, mx = par[:4]
        normalisation = (erf((mx - mu) / 2 ** 0.5 / sigma) - erf((mn - mu) / 2 ** 0.5 / sigma)) / 2
        return erfinv(2 * u * normalisation + erf((mn - mu) / 2 ** 0.5 / sigma)) * 2 ** 0.5 * sigma + mu
    if kind == "PowerLaw":
        alpha, mn, mx = par[:3]
        if alpha == -1:
            return mn * np.exp(u * np.log(mx / mn))
        return (mn ** (1 + alpha) + u * (mx ** (1 + alpha) - mn ** (1 + alpha))) ** (1.0 / (1 + alpha))
    if kind == "Triangular":      # scipy.stats.triang.ppf form of bilby's Triangular
        mode, mn, mx = par[:3]
        fc = (mode - mn) / (mx - mn)
        lo = mn + np.sqrt(np.maximum(u, 0) * (mx - mn) * (mode - mn))
        hi = mx - np.sqrt(np.maximum(1 - u, 0) * (mx - mn) * (mx - mode))
        return np.where(u < fc, lo, hi)
    if kind == "Interped":
        cdf, grid = table
        return np.interp(u, cdf, grid)
    raise ValueError(kind)
