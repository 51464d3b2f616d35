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
