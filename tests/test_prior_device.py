"""Priors on the device: ``nmma_b200_set_priors / prior_transform / prior_sample /
logl_sweep`` against the NumPy restatement in ``oracle/philox.py``.

Bars: the unit-cube draws (Philox4x32-10 integer arithmetic) are bit-exact; the transform is fp64 with
libm-vs-CUDA differences of a few ulp in acos / asin / erfinv / exp / pow, tolerance 1e-13 relative (written below);
the sweep equals ``log_likelihood_batch`` of the same points bit for bit.
"""
import numpy as np
import pytest

from helpers import SENTINEL, build_pair

KINDS = ["Uniform", "DeltaFunction", "Sine", "Cosine", "Gaussian", "TruncatedGaussian", "PowerLaw", "Triangular",
         "Interped"]


# ---------------------------------------------------------------------------------------------------
# CPU: the checker itself
# ---------------------------------------------------------------------------------------------------
def test_philox_known_answer_vectors():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    from oracle import philox as ph
    kats = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
            ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
            ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
             [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kats:
        got = ph.philox4x32_10(*[np.array([c], np.uint64) for c in ctr], *key)
        assert [int(g[0]) for g in got] == want


def test_unit_cube_is_counter_based_and_uniform():
    from oracle import philox as ph
    u = ph.unit_cube(7, 0, 20000, 5)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1 / 12) < 2e-3
    assert np.abs(np.corrcoef(u.T) - np.eye(5)).max() < 0.03
    # any split of the index range reproduces the same draws
    assert np.array_equal(u[5000:9000], ph.unit_cube(7, 5000, 4000, 5))
    assert not np.array_equal(u, ph.unit_cube(8, 0, 20000, 5))


def _zoo():
    """One column per prior kind: (kind name, params[4], table)."""
    from nmma_b200.core import priors as pr
    objs = [pr.Uniform(-2.0, 0.1), pr.DeltaFunction(0.37), pr.Sine(0.0, np.pi / 2), pr.Cosine(-0.3, 1.2),
            pr.Gaussian(40.0, 3.0), pr.TruncatedGaussian(1.0, 0.5, 0.2, 1.4), pr.PowerLaw(2.0, 1.0, 200.0),
            pr.LogUniform(1e-3, 2.0), pr.Triangular(0.3, 0.0, 1.0), pr.Interped([0.0, 0.5], [4.0, 0.0], 0.0, 0.5)]
    names = [f"p{i}" for i in range(len(objs))]
    return pr.PriorDict(dict(zip(names, objs))), names


def test_host_priors_equal_oracle_rescale():
    """nmma_b200.core.priors (the bilby stand-in) and the oracle restatement agree to the last bit or two."""
    from oracle import philox as ph
    pd, names = _zoo()
    kinds, params, tables = pd.device_plan(names)
    assert len(kinds) == len(names) and params.shape == (len(names), 4)
    u = ph.unit_cube(3, 0, 4096, len(names))
    kind_names = {0: "Uniform", 1: "DeltaFunction", 2: "Sine", 3: "Cosine", 4: "Gaussian", 5: "TruncatedGaussian",
                  6: "PowerLaw", 7: "Triangular", 8: "Interped"}
    for j, n in enumerate(names):
        want = ph.rescale_column(kind_names[int(kinds[j])], params[j], u[:, j], tables.get(j))
        got = np.asarray(pd[n].rescale(u[:, j]), float) * np.ones_like(u[:, j])
        np.testing.assert_allclose(got, want, rtol=1e-15, atol=0)
        assert np.isfinite(want).all()


def test_reference_prior_files_have_device_plans():
    """Every kilonova prior file the north star names maps onto device prior kinds."""
    from nmma_b200 import synthetic as syn
    pd = syn.bu2019lm_prior()
    kinds, params, tables = pd.device_plan()
    assert len(kinds) == 6 and not tables
    assert sorted(set(int(k) for k in kinds)) == [0, 2]          # Uniform + Sine


# ---------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _zoo_engine():
    from nmma_b200.engine import KilonovaEngine
    pd, names = _zoo()
    eng = KilonovaEngine(0)
    eng.set_priors(*pd.device_plan(names))
    return eng, pd, names


@pytest.mark.gpu
def test_device_transform_matches_oracle(torch_cuda):
    from oracle import philox as ph
    eng, pd, names = _zoo_engine()
    kinds, params, tables = pd.device_plan(names)
    kind_names = dict(enumerate(KINDS))
    rng = np.random.default_rng(11)
    u = rng.uniform(size=(50000, len(names)))
    u[0], u[1], u[2] = 0.0, 1.0, 0.5                              # the ends of the unit interval
    u[3] = np.nextafter(1.0, 0.0)
    got = eng.prior_transform(u).cpu().numpy()
    for j in range(len(names)):
        want = ph.rescale_column(kind_names[int(kinds[j])], params[j], u[:, j], tables.get(j))
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got[:, j]), fin), names[j]
        assert np.array_equal(got[~fin, j], want[~fin]) or not (~fin).any()      # +-inf at u = 0 / 1 (Gaussian)
        scale = np.maximum(np.abs(want[fin]), 1e-3)
        err = np.abs(got[fin, j] - want[fin]) / scale
        # erfinv near u -> 0/1 is ill-conditioned (d erfinv / du ~ 1 / pdf): compare away from the last 1e-9
        core = (u[fin, j] > 1e-9) & (u[fin, j] < 1 - 1e-9)
        assert err[core].max() < 1e-13, (KINDS[int(kinds[j])], err[core].max())
    # in-place transform (unit and points alias)
    t = torch_cuda.from_numpy(u).cuda()
    eng.prior_transform(t, out=t)
    assert np.array_equal(t.cpu().numpy(), got, equal_nan=True)
    # linear priors are bit-exact (same fp64 operations)
    for j in (0, 1):
        want = ph.rescale_column(kind_names[int(kinds[j])], params[j], u[:, j])
        assert np.array_equal(got[:, j], want)
    eng.close()


@pytest.mark.gpu
def test_device_sampler_is_bit_exact_philox(torch_cuda):
    from oracle import philox as ph
    eng, pd, names = _zoo_engine()
    P = len(names)
    seed = 0x1234_5678_9ABC_DEF0
    pts, unit = eng.prior_sample(30000, seed=seed, first_index=0, return_unit=True)
    unit = unit.cpu().numpy()
    assert np.array_equal(unit, ph.unit_cube(seed, 0, 30000, P))
    # a 64-bit first index and an odd split reproduce the same stream (sharding invariance)
    a = eng.prior_sample(1000, seed=seed, first_index=12345).cpu().numpy()
    assert np.array_equal(a, pts[12345:13345].cpu().numpy(), equal_nan=True)
    big = (1 << 33) + 17
    _, ub = eng.prior_sample(257, seed=seed, first_index=big, return_unit=True)
    assert np.array_equal(ub.cpu().numpy(), ph.unit_cube(seed, big, 257, P))
    # transform of the drawn cube == draws
    assert np.array_equal(eng.prior_transform(unit).cpu().numpy(), pts.cpu().numpy(), equal_nan=True)
    eng.close()


@pytest.mark.gpu
def test_sweep_equals_batch_on_drawn_points(torch_cuda):
    """BASELINE.json configs[1]/[4] shape: prior draws on the device, log L without host traffic."""
    from nmma_b200 import synthetic as syn
    from oracle import harness, philox as ph
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    core = syn.random_model("Bu2019lm", filters, seed=0)
    priors = syn.bu2019lm_prior()
    lik, olik, fixed, cols = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors)
    n = 80000
    logl, pts = lik.log_likelihood_sweep(n, seed=99, return_points=True)
    assert pts.shape == (n, len(cols))
    again = lik.log_likelihood_batch(pts, cols)
    assert torch_cuda.equal(logl, again)
    # without the points buffer (scratch blocks), and split as two ranks would
    assert torch_cuda.equal(lik.log_likelihood_sweep(n, seed=99), logl)
    # (both halves above the tensor-core threshold: the same kernel, a point's value does not depend on its tile)
    lo = lik.log_likelihood_sweep(40001, seed=99, first_index=0)
    hi = lik.log_likelihood_sweep(n - 40001, seed=99, first_index=40001)
    assert torch_cuda.equal(torch_cuda.cat([lo, hi]), logl)
    # a short block takes the FFMA kernel: same draws, log L equal to fp32 rounding of the surrogate
    short = lik.log_likelihood_sweep(3000, seed=99, first_index=1000)
    ok = logl[1000:4000] != SENTINEL
    assert torch_cuda.equal(short != SENTINEL, ok)
    rel = ((short - logl[1000:4000]).abs() / logl[1000:4000].abs().clamp(min=1.0))[ok]
    assert float(rel.max()) < 1e-4
    # the draws are the prior's: unit cube from the oracle generator, rescaled by the host PriorDict
    unit = ph.unit_cube(99, 0, 512, len(cols))
    want = np.stack([np.asarray(priors[k].rescale(unit[:, i]), float) for i, k in enumerate(cols)], axis=1)
    np.testing.assert_allclose(pts[:512].cpu().numpy(), want, rtol=1e-13, atol=0)
    # and the oracle agrees on the first points
    ref = harness.oracle_logl(olik, fixed, pts[:48].cpu().numpy(), cols)
    got = logl[:48].cpu().numpy()
    assert np.array_equal(got == SENTINEL, ref == SENTINEL)
    assert (np.abs(got - ref) / np.maximum(1, np.abs(ref))).max() < 1e-4
    # ultranest-style vectorised callables
    transform, loglike = lik.vectorized()
    theta = transform(unit)
    np.testing.assert_allclose(theta, want, rtol=1e-13, atol=0)
    assert np.array_equal(loglike(theta), lik.log_likelihood_batch(theta, cols))


@pytest.mark.gpu
def test_prior_error_paths(torch_cuda):
    from nmma_b200 import _lib as L
    from nmma_b200.engine import KilonovaEngine
    eng = KilonovaEngine(0)
    with pytest.raises(L.NmmaB200Error) as e:
        eng.prior_P = 2
        eng.prior_sample(4)
    assert e.value.code == L.ERR_STATE
    for kinds, par in (([L.PR_UNIFORM], [[1.0, 0.0, 0, 0]]), ([L.PR_GAUSSIAN], [[0.0, -1.0, 0, 0]]),
                       ([L.PR_POWERLAW], [[2.0, 0.0, 1.0, 0]]), ([17], [[0, 0, 0, 0]]),
                       ([L.PR_INTERPED], [[0, 0, 0, 0]])):
        with pytest.raises(L.NmmaB200Error) as e:
            eng.set_priors(kinds, par)
        assert e.value.code == L.ERR_ARG
    eng.set_priors([L.PR_UNIFORM, L.PR_SINE], [[0, 1, 0, 0], [0, np.pi, 0, 0]])
    assert eng.prior_sample(0).shape == (0, 2)
    eng.close()
