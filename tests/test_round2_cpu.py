"""CPU tests added in round 2: the dynesty pool seam with a stub engine, the extinction oracle's pins,
the effective-wavelength table, Constraint / Ebv layouts, and the prior construction of ``create_prior_from_args``."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import REFERENCE, ROOT
from helpers import fixture_core, synthetic_observations


# ---- BatchPool.map ----------------------------------------------------------------------------------------------
class _StubLikelihood:
    """Counts calls; log L = sum of the row, prior transform = 2 u (stands in for the GPU engine)."""

    def __init__(self, columns):
        self.columns = list(columns)
        self.batch_calls = 0
        self.transform_calls = 0
        self.sub_model = object()

    def log_likelihood_batch(self, pts, columns=None, out=None):
        self.batch_calls += 1
        return np.asarray(pts, float).sum(axis=1)

    def log_likelihood(self, parameters):
        return float(sum(parameters[k] for k in self.columns))

    def prior_transform_batch(self, unit, columns=None):
        self.transform_calls += 1
        arr = 2.0 * np.asarray(unit, float)
        return SimpleNamespace(cpu=lambda: SimpleNamespace(numpy=lambda: arr))


def test_batch_pool_dispatches_on_the_function_not_on_the_shape():
    from nmma_b200.em.em_likelihood import BatchPool
    lik = _StubLikelihood(["a", "b", "c"])
    pool = BatchPool(lik)
    pts = np.arange(12.0).reshape(4, 3)
    # likelihood calls: array thetas, dict thetas, bound method of the likelihood, sampler-style wrapper objects
    assert pool.map(pool.loglike, list(pts)) == [3.0, 12.0, 21.0, 30.0] and lik.batch_calls == 1
    assert pool.map(lik.log_likelihood, [dict(zip("abc", r)) for r in pts]) == [3.0, 12.0, 21.0, 30.0] and lik.batch_calls == 2

    class FunctionWrapper:          # dynesty.utils._function_wrapper keeps the user function in .func
        def __init__(self, func):
            self.func = func

        def __call__(self, x):
            return self.func(x)

    class LogLikelihood:            # dynesty.utils._LogLikelihood keeps it in .loglikelihood
        def __init__(self, f):
            self.loglikelihood = f

    assert pool.map(FunctionWrapper(pool.loglike), list(pts)) == [3.0, 12.0, 21.0, 30.0] and lik.batch_calls == 3
    assert pool.map(LogLikelihood(FunctionWrapper(pool.loglike)), list(pts)) == [3.0, 12.0, 21.0, 30.0] and lik.batch_calls == 4
    # prior transform with the SAME [N, P] shape must return transformed points, never log-likelihoods
    u = np.full((5, 3), 0.25)
    out = pool.map(pool.prior_transform, list(u))
    assert lik.transform_calls == 1 and lik.batch_calls == 4 and np.array_equal(np.array(out), np.full((5, 3), 0.5))
    user_pt = lambda x: np.asarray(x) + 1.0     # noqa: E731  a user's own prior_transform: mapped verbatim
    out = pool.map(user_pt, list(u))
    assert lik.batch_calls == 4 and np.array_equal(np.array(out), np.full((5, 3), 1.25))
    # object arguments (dynesty's evolve_point receives SamplerArgument tuples): no array conversion, no TypeError
    args = [SimpleNamespace(u=np.zeros(3), scale=1.0), SimpleNamespace(u=np.ones(3), scale=2.0)]
    assert pool.map(lambda a: a.scale, args) == [1.0, 2.0]
    assert pool.map(pool.loglike, []) == []
    # ragged / wrong-width input to the likelihood falls back to per-point calls instead of raising inside np.asarray
    assert pool.map(pool.loglike, [np.arange(3.0)]) == [3.0]
    assert pool.loglike({"a": 1.0, "b": 2.0, "c": 3.0}) == 6.0
    with pool as p:
        assert p is pool


# ---- extinction -------------------------------------------------------------------------------------------------
def test_extinction_oracle_pins():
    """What can be pinned without dust_extinction (oracle/extinction.py header): Pei's own normalisation, the SMC shape,
    the reference's validity window and the Ebv = 0 / out-of-range identities."""
    from oracle import extinction as X
    # Pei (1992): xi = A_lambda / A_B is normalised at the B band; his analytic fit reproduces 1 to a few per cent there,
    # and dust_extinction's A(V) reference (x AbAv) gives A(0.55 um) / A(V) = 1 to the same accuracy
    assert X.p92_smc_axav(0.44) / X.P92_ABAV == pytest.approx(1.0, abs=0.05)
    assert X.p92_smc_axav(0.55) == pytest.approx(1.0, abs=0.05)
    assert X.P92_ABAV == pytest.approx(1.3247, abs=1e-4)
    lam = np.geomspace(0.1, 3.0, 200)
    ax = X.p92_smc_axav(lam)
    assert np.all(np.diff(ax) < 0)                                   # SMC: monotonic, no 2175 A bump
    assert 4.0 < X.p92_smc_axav(0.15) < 5.5                          # steep far-UV rise (Gordon+ 2003 SMC bar: ~4.5-5)
    nu = 299792458.0 / np.array([4866.46e-10, 21656e-10, 1e-12, 1.0])   # g, Ks, gamma rays (nu > 2e16 Hz), 1 m radio (below range)
    ext = X.extinction_factor_p92_smc(nu, 0.3, 0.01)
    assert ext[2] == 1.0 and ext[3] == 1.0 and 0 < ext[0] < ext[1] < 1
    assert X.get_extinction_mags(nu, 0.0, 0.01).tolist() == [0.0] * 4
    m = X.get_extinction_mags(nu, 0.3, 0.01)
    assert m[0] == pytest.approx(X.p92_smc_axav(0.486646 / 1.01) * 2.93 * 0.3, rel=1e-12) and m[2] == 0.0
    # redshift moves the filter blue-ward in the host frame: more extinction
    assert X.get_extinction_mags(nu[:1], 0.3, 0.5)[0] > m[0]
    assert X.get_extinction_mags(nu[:2], 0.2, 0.0, "G23_MW", [3.0, 0.3]).tolist() == pytest.approx([0.6, 0.06])


def test_wave_eff_table_reproduces_reference_constants():
    """tools/make_wave_eff.py restates sncosmo's wave_eff; the reference hard-codes six PS1 values that came from sncosmo
    (``lambdas_sloan``, nmma/em/utils.py:712-714): g 4866.46, r 6214.6, open 7127.0, i 7544.6, z 8679.5, y 9633.3."""
    from nmma_b200.em.utils import get_default_filts_lambdas
    tab = json.load(open(os.path.join(ROOT, "nmma_b200", "data", "wave_eff.json")))["wave_eff"]
    for name, want, tol in [("ps1::g", 4866.46, 0.01), ("ps1::r", 6214.6, 0.06), ("ps1::open", 7127.0, 0.2),
                            ("ps1::i", 7544.6, 0.06), ("ps1::z", 8679.5, 0.06), ("ps1::y", 9633.3, 0.06)]:
        assert tab[name] == pytest.approx(want, abs=tol), name
    filts, lambdas = get_default_filts_lambdas(["ps1::g", "g", "2massks", "radio-3GHz", "X-ray-1keV", "radio-10GHz", "nonsense"])
    assert filts == ["ps1::g", "g", "2massks", "radio-3GHz", "X-ray-1keV", "radio-10GHz"]
    assert lambdas[0] == pytest.approx(4866.46e-10, rel=1e-5) and lambdas[1] == 4866.46e-10
    assert lambdas[3] == pytest.approx(0.0999308, rel=1e-6) and lambdas[5] == pytest.approx(0.0299792458, rel=1e-9)
    assert lambdas[4] == pytest.approx(1.2398e-9, rel=1e-4)


@pytest.mark.reference
def test_wave_eff_table_is_current():
    """The committed table equals what tools/make_wave_eff.py produces from the vendored transmission curves."""
    import subprocess
    import sys
    src = os.path.join(REFERENCE, "nmma-data", "sncosmo", "bandpasses")
    if not os.path.isdir(src):
        pytest.skip("nmma-data is not checked out")
    before = open(os.path.join(ROOT, "nmma_b200", "data", "wave_eff.json")).read()
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_wave_eff.py"), src], check=True, capture_output=True)
    assert open(os.path.join(ROOT, "nmma_b200", "data", "wave_eff.json")).read() == before


# ---- layouts ------------------------------------------------------------------------------------------------------
def _lik(priors, extinction_law=None):
    from nmma_b200.em import EMTransientLikelihood, FilterSystematicsHandler, SVDLightCurveModel
    filters = ["ztfr", "sdssu", "2massks"]
    core = fixture_core("mlp", filters)
    model = SVDLightCurveModel("Bu2019nsbh", svd_mag_model=core, interpolation_type="tensorflow", filters=filters,
                               extinction_law=extinction_law)
    lc = synthetic_observations(filters, np.random.default_rng(0))
    return EMTransientLikelihood(model, lc, FilterSystematicsHandler(filters, None, 0.5, lc[0]), priors, filters=filters)


def test_constraint_and_extinction_layouts():
    from nmma_b200 import _lib as L
    from nmma_b200.core.priors import Constraint, PriorDict, Sine, Uniform
    priors = PriorDict()
    priors["luminosity_distance"] = Uniform(10.0, 200.0)
    priors["inclination_EM"] = Sine(0.0, np.pi / 2)
    priors["log10_mej_dyn"] = Uniform(-2.0, -1.05)
    priors["log10_mej_wind"] = -1.5
    priors["KNtheta"] = Constraint(10.0, 80.0)
    priors["luminosity_distance_c"] = Constraint(0.0, 1.0)
    lik = _lik(priors)
    assert lik.columns == ["luminosity_distance", "inclination_EM", "log10_mej_dyn"] and sorted(lik.constraints) == ["KNtheta", "luminosity_distance_c"]
    with pytest.raises(NotImplementedError, match="conversion chain does not produce"):
        lik.sub_model.plan_layout(lik.columns)
    del priors["luminosity_distance_c"]
    lik = _lik(priors)
    plan = lik.sub_model.plan_layout(lik.columns)
    srcs, lo, hi = plan["constraints"]
    assert (srcs[0].col, srcs[0].transform) == (1, L.XF_RAD2DEG) and lo == [10.0] and hi == [80.0]
    assert plan["ext"] is None
    priors["Ebv"] = Uniform(0.0, 0.5)
    lik = _lik(priors)
    law, src, nu0, coef = lik.sub_model.plan_layout(lik.columns)["ext"]
    assert law == L.EXT_P92_SMC_HOST and src.col == 3 and coef is None
    assert nu0 == pytest.approx(299792458.0 / (1e-10 * np.array([6421.80, 3594.33, 21656.09])), rel=1e-5)
    lik = _lik(priors, extinction_law="G23_MW")
    with pytest.raises(NotImplementedError, match="dust_extinction"):
        lik.sub_model.plan_layout(lik.columns)
    lik.sub_model.light_curve_model.extinction_coefficients = {"ztfr": 2.6, "sdssu": 4.8}
    law, src, nu0, coef = lik.sub_model.plan_layout(lik.columns)["ext"]
    assert law == L.EXT_LINEAR and coef.tolist() == [2.6, 4.8, 0.0]
    priors["Ebv"] = 0.0                                                   # DeltaFunction(0): the reference skips the correction
    assert _lik(priors).sub_model.plan_layout(["luminosity_distance", "inclination_EM", "log10_mej_dyn"])["ext"] is None
    with pytest.raises(ValueError, match="Unknown extinction_law"):
        priors["Ebv"] = 0.1
        _lik(priors, extinction_law="CCM89").sub_model.plan_layout(["luminosity_distance", "inclination_EM", "log10_mej_dyn"])


def test_dict_columns_follow_the_dict():
    """log_likelihood(dict) evaluates the dict as given: which keys become extra columns."""
    from nmma_b200.core.priors import PriorDict, Sine, Uniform
    priors = PriorDict()
    priors["luminosity_distance"] = Uniform(10.0, 200.0)
    priors["inclination_EM"] = Sine(0.0, np.pi / 2)
    priors["log10_mej_dyn"] = Uniform(-2.0, -1.05)
    priors["log10_mej_wind"] = -1.5
    priors["timeshift"] = 0.0
    sm = _lik(priors).sub_model
    sm._dict_base = sm.default_columns()
    base = {"luminosity_distance": 40.0, "inclination_EM": 0.3, "log10_mej_dyn": -1.5, "log10_mej_wind": -1.5, "timeshift": 0.0}
    assert sm._dict_columns(base) == sm._dict_base                                        # the sampler's own dict: hot path
    assert sm._dict_columns(dict(base, KNtheta=17.2, redshift=0.01, Ebv=0.0, geocent_time=1.2e9, label="x")) == sm._dict_base
    assert sm._dict_columns(dict(base, timeshift=-0.4)) == sm._dict_base + ["timeshift"]  # fixed key, different value
    assert sm._dict_columns(dict(base, Ebv=0.2, log10_mej_wind=-1.2)) == sm._dict_base + ["log10_mej_wind", "Ebv"]


def test_create_prior_from_args(tmp_path):
    from nmma_b200.em.prior import create_prior_from_args
    from nmma_b200.em.systematics import FilterSystematicsHandler
    pf = tmp_path / "p.prior"
    pf.write_text("luminosity_distance = Uniform(minimum=1, maximum=200, name='luminosity_distance')\ntimeshift = 0.\n")
    args = SimpleNamespace(prior=str(pf), use_Ebv=False, Ebv_max=0.5724)
    p = create_prior_from_args(args, None)
    assert list(p) == ["luminosity_distance", "timeshift", "Ebv"] and p["Ebv"].peak == 0.0
    args.use_Ebv = True
    p = create_prior_from_args(args, None)
    assert p["Ebv"].__class__.__name__ == "Interped" and p["Ebv"].maximum == 0.5724
    u = np.linspace(0, 1, 11)
    assert np.allclose(p["Ebv"].rescale(u), 0.5724 * (1 - np.sqrt(1 - u)), atol=2e-4)     # inverse CDF of the triangular density
    yaml = {"config": {"withTime": {"value": False, "filters": [None], "time_nodes": 4, "type": "Uniform", "minimum": 0, "maximum": 2},
                       "withoutTime": {"value": True, "type": "Uniform", "minimum": 0, "maximum": 2}}}
    h = FilterSystematicsHandler(["ztfr"], yaml, 1.0, {"ztfr": np.array([1.0, 2.0])})
    p = create_prior_from_args(args, h)
    assert "em_syserr" in p
    for bad in (dict(fetch_Ebv_from_dustmap=True), dict(Hubble_weight="w.dat"), dict(conditional_gaussian_prior_thetaObs=True),
                dict(fits_file="x.fits")):
        with pytest.raises(NotImplementedError):
            create_prior_from_args(SimpleNamespace(prior=str(pf), use_Ebv=False, Ebv_max=0.5, **bad), None)


def test_fast_log_ndtr_restatement():
    """csrc/kernels.cuh: fast_log_ndtr (log Phi(b) of the FAST back end's finite-detection-limit class) restated operation
    by operation in fp32 NumPy and checked against scipy.special.log_ndtr: pins the Chebyshev coefficients (Numerical
    Recipes erfcc, taken in the log domain), the 1e-3 switch and the 8.3 cut-off."""
    from scipy.special import log_ndtr
    f = np.float32
    b = np.linspace(-40.0, 9.0, 200001).astype(f)
    y = np.abs(b) * f(0.70710678)
    t = (f(1) / (f(0.5) * y + f(1))).astype(f)
    coef = [0.17087277, -0.82215223, 1.48851587, -1.13520398, 0.27886807, -0.18628806, 0.09678418, 0.37409196, 1.00002368,
            -1.26551223]
    p = f(coef[0])
    for c in coef[1:]:
        p = (p * t + f(c)).astype(f)
    lg = (np.log((f(0.5) * t).astype(f)).astype(f) + (p - y * y).astype(f)).astype(f)
    q = np.exp(lg).astype(f)
    pos = np.where(q < 1e-3, -q * (1 + f(0.5) * q), np.log((1 - q).astype(f))).astype(f)
    got = np.where(b > 0, np.where(b > 8.3, 0.0, pos), lg)
    ref = log_ndtr(b.astype(np.float64))
    assert (np.abs(got - ref) / np.maximum(1.0, np.abs(ref))).max() < 6e-7


class _StubModel:
    """Minimal LightCurveModel for the combined-container tests: constant magnitudes per filter on its own time grid."""

    def __init__(self, name, filters, times, offset):
        self.model, self.filters, self.model_times, self.model_parameters = name, list(filters), np.asarray(times, float), ["a"]
        self.offset, self.citation, self.good_parameters = offset, {name: "cite"}, True

    def parameter_conversion(self, p):
        return p

    def check_vs_priors(self, p):
        pass

    def generate_lightcurve(self, t, p):
        return {f: 20.0 + self.offset + 0.1 * i + 0.05 * np.asarray(t, float) for i, f in enumerate(self.filters)}

    def gen_detector_lc(self, p, sample_times=None):
        t = self.model_times * 1.01 + self.offset
        return t, {f: 20.0 + self.offset + 0.1 * i + 0.05 * t for i, f in enumerate(self.filters)}


def test_combined_model_container_stacks_fluxes():
    """CombinedLightCurveModelContainer (nmma/em/model.py:1342-1510): flux sum per filter, union time grid, +inf outside a
    sub-model's range; compared with the reference's own class when /root/reference is available (this container)."""
    from nmma_b200.em import CombinedLightCurveModelContainer
    a = _StubModel("A", ["ps1::g", "ps1::r"], np.linspace(0, 10, 5), 0.0)
    b = _StubModel("B", ["ps1::r", "sdssu"], np.linspace(0, 8, 4), 1.0)
    comb = CombinedLightCurveModelContainer([a, b])
    t = np.linspace(0, 5, 4)
    lc = comb.generate_lightcurve(t, {"a": 1.0})
    ma, mb = a.generate_lightcurve(t, {})["ps1::r"], b.generate_lightcurve(t, {})["ps1::r"]
    assert np.allclose(lc["ps1::r"], -2.5 * np.log10(10 ** (-0.4 * ma) + 10 ** (-0.4 * mb)), rtol=0, atol=1e-12)
    assert np.allclose(lc["ps1::g"], a.generate_lightcurve(t, {})["ps1::g"], rtol=0, atol=1e-12) and set(lc) == {"ps1::g", "ps1::r", "sdssu"}
    tj, lcj = comb.gen_detector_lc({"a": 1.0})
    assert np.array_equal(tj, np.array(sorted(set((a.model_times * 1.01).tolist()) | set((b.model_times * 1.01 + 1.0).tolist()))))
    assert lcj["ps1::r"][0] == pytest.approx(20.1, abs=1e-12)                       # t = 0: model B has not started (+inf -> no flux)
    assert comb.model_parameters == ["a", "a"] and comb.good_parameters
    with pytest.raises(NotImplementedError):
        comb.new_engine()                                               # no device engine: the batched likelihood is single-model
    if not os.path.isdir("/root/reference/nmma"):
        return
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import reference_stubs as RS
    ref_cls = RS.load_reference()["model"].CombinedLightCurveModelContainer
    rc = ref_cls([a, b])
    rlc = rc.generate_lightcurve(t, {"a": 1.0})
    for f in lc:
        assert np.array_equal(lc[f], rlc[f]), f
    rt, rlcj = rc.gen_detector_lc({"a": 1.0})
    assert np.array_equal(rt, tj)
    for f in lcj:
        assert np.array_equal(lcj[f], rlcj[f]), f


def test_fp16_split_operand_scheme_emulated():
    """CPU emulation (tools/tc_numerics.py: mlp_f16x3) of the tensor-core front end's operand scheme -- fp16 hi/lo pairs, exact
    power-of-two scalings of the rows of [W1; b1], of the columns of W2 and of every point, h_hi rounded toward zero with the
    ReLU -- on the trained Bu2019nsbh fixture weights: as close to the fp64 network as NumPy's fp32, and the scaled
    activations stay below 2^14 (the bound csrc/tc_kernel.cuh relies on) for inputs far outside the training range."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import tc_numerics as T
    z = np.load(os.path.join(ROOT, "tests", "golden", "bu2019nsbh_fixture.npz"), allow_pickle=True)
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-0.1, 1.5, size=(512, 3)), 10.0 ** rng.uniform(-4, 4, size=(256, 3)) * rng.choice([-1, 1], (256, 3))])
    for f in ("ztfr", "sdssu", "2massks"):
        W1, b1, W2, b2 = (z[f"{f}/{n}"] for n in ("W1", "b1", "W2", "b2"))
        xa = np.concatenate([x.astype(np.float32), np.ones((len(x), 1), np.float32)], 1)
        Wa = np.concatenate([W1, b1[None, :]], 0).astype(np.float32)
        h = np.maximum(xa.astype(np.float64) @ Wa.astype(np.float64), 0)
        exact = h @ W2.astype(np.float64) + b2
        scale = (np.abs(xa.astype(np.float64)) @ np.abs(Wa.astype(np.float64))) @ np.abs(W2.astype(np.float64)) + np.abs(b2)
        got = T.mlp_f16x3(x, W1, b1, W2, b2)                       # asserts 2^e relu(v) < 2^14 inside
        fp32 = T.mm32(np.maximum(T.mm32(xa, Wa), 0), W2) + b2
        e_split = (np.abs(got - exact) / scale).max()
        e_fp32 = (np.abs(fp32 - exact) / scale).max()
        assert e_split < 2e-7 and e_split < 4 * e_fp32 + 2e-8, (f, e_split, e_fp32)
