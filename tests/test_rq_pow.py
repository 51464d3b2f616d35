"""The table + series pow of the GP front end (csrc/kernels.cuh: rq_pow) restated operation by operation in NumPy fp64
and checked against np.power: pins the constants, table geometry and polynomial degrees; the device code itself is
covered by the GP parity tests (tests/test_gpu_parity.py::test_ka2017_gp, ::test_fixture_mags_and_logl[gp])."""
import numpy as np

NT = 256
_C = 1.0 + (np.arange(NT) + 0.5) / NT
INV_C, L2C, E2 = 1.0 / _C, np.log2(_C), 2.0 ** (np.arange(NT) / NT)


def rq_pow(base, alpha):
    bits = base.view(np.int64)
    e = ((bits >> 52) & 0x7FF) - 1023
    i = (bits >> 44) & (NT - 1)
    m = ((bits & 0x000FFFFFFFFFFFFF) | 0x3FF0000000000000).view(np.float64)
    u = m * INV_C[i] - 1.0
    p = -1.0 / 6.0
    for c in (0.2, -0.25, 1.0 / 3.0, -0.5, 1.0):
        p = p * u + c
    t2 = -alpha * ((p * u) * 1.4426950408889634 + (e + L2C[i]))
    k = np.rint(t2 * NT).astype(np.int64)
    x = (t2 - k / NT) * 0.6931471805599453
    q = 1.0 / 120.0
    for c in (1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0):
        q = q * x + c
    return np.ldexp(E2[k & (NT - 1)] * q, (k >> 8).astype(np.int64))


def test_rq_pow_restatement_accuracy():
    rng = np.random.default_rng(0)
    base = 1.0 + 10.0 ** rng.uniform(-9, 4, 400_000)          # 1 + r^2 / (2 alpha l^2)
    alpha = 10.0 ** rng.uniform(-3, 0.7, 400_000)
    ref = np.power(base, -alpha)
    ok = ref > 1e-300
    err = np.abs(rq_pow(base, alpha)[ok] - ref[ok]) / ref[ok]
    assert err.max() < 2e-14, err.max()
    # base = 1 (a training point hit exactly) and monotone decrease in base, both to rounding
    assert np.abs(rq_pow(np.ones(4), np.array([0.01, 0.3, 1.0, 4.0])) - 1.0).max() < 2e-15
    b = np.linspace(1.0, 50.0, 20001)
    assert np.all(np.diff(rq_pow(b, np.full_like(b, 0.37))) <= 1e-15)


# ---- fused GP kernel (csrc/gp_kernel.cuh: gf_pow): fp32-reciprocal cell centres, cubic log2, 1.5 * 2^52 rounding trick ----
_CF = (np.float32(1.0) + (np.arange(NT, dtype=np.float32) + np.float32(0.5)) / np.float32(NT))
_R = (np.float32(1.0) / _CF).astype(np.float64)          # the table holds the log of exactly this fp32 value
_L2R = -np.log2(_R)
_MAGIC = 6755399441055744.0


def gf_pow(r2, q, alpha):
    na = -256.0 * alpha
    base = r2 * q + 1.0
    bits = base.view(np.int64)
    hi = (bits >> 32).astype(np.int64)
    i = (hi >> 12) & (NT - 1)
    e = ((hi >> 20) & 0x7FF) - 1023
    rs = np.ldexp(_R[i], -e.astype(np.int64))              # exponent taken off r_i (integer add on its high word)
    u = base * rs - 1.0                                      # an fma on the device (exact); here |error| < 2^-53
    p = -0.36067471452205946 * u + 0.4808994921226281
    p = p * u - 0.7213475204440083
    p = p * u + 1.4426950408883954
    lg2 = (p * u + _L2R[i]) + e
    t = na * lg2
    s = t + _MAGIC
    k = np.maximum((s.view(np.int64) & 0xFFFFFFFF).astype(np.uint32).view(np.int32).astype(np.int64), -1020 * 256)
    xr = t - (s - _MAGIC)
    g = 2.2393953277407236e-12 * xr + 3.308302983832675e-9
    g = g * xr + 3.665565596910102e-6
    g = (g * xr + 0.0027076061740622769) * xr
    e2 = E2[k & (NT - 1)]
    return np.ldexp(e2 * g + e2, (k >> 8).astype(np.int64))


def test_gf_pow_restatement_accuracy():
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(1)
    n = 200_000
    r2 = 10.0 ** rng.uniform(-12, 1.5, n)
    q = 10.0 ** rng.uniform(-3, 4, n)
    alpha = 10.0 ** rng.uniform(-3, 0.7, n)
    got = gf_pow(r2, q, alpha)
    ref = np.power(r2 * q + 1.0, -alpha)
    ok = ref > 1e-300
    assert (np.abs(got[ok] - ref[ok]) / ref[ok]).max() < 2e-14
    # against a 40-digit reference (np.power itself is only good to ~1e-16 relative): a subset
    idx = rng.choice(n, 2000, replace=False)
    worst = 0.0
    for j in idx:
        b = mp.mpf(float(r2[j])) * mp.mpf(float(q[j])) + 1        # the device rounds base once: allow that
        exact = mp.power(mp.mpf(float(np.float64(r2[j] * q[j] + 1.0))), -mp.mpf(float(alpha[j])))
        worst = max(worst, float(abs(mp.mpf(float(got[j])) - exact) / exact))
    assert worst < 1e-14, worst
    # base = 1 (a training point hit exactly), large alpha inside the sklearn bound, underflow to ~0, monotone decrease
    assert np.abs(gf_pow(np.zeros(4), np.ones(4), np.array([0.01, 0.3, 1.0, 4.0])) - 1.0).max() < 5e-15   # alpha ln2 x 1.1e-15 (cubic)
    big = gf_pow(np.array([3.0, 1e6]), np.array([1.0, 1.0]), np.array([1e5, 1e5]))
    assert np.all(big >= 0.0) and np.all(big < 1e-300)
    b = np.linspace(0.0, 49.0, 20001)
    assert np.all(np.diff(gf_pow(b, np.ones_like(b), np.full_like(b, 0.37))) <= 1e-15)
