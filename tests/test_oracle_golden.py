"""The oracle pinned against the reference's own known-answer test and fixtures (CPU only)."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, fixture_core, oracle_ready_core
from conftest import REFERENCE


def _golden_mag(core):
    """nmma/tests/joint_analysis_pipeline.py:108-120: Bu2019nsbh / tensorflow / ztfr, em_tmin=0.1,
    em_tmax=10, em_tstep=0.5, injection row 0, error budget 0.1, generation_seed=42 -> data['ztfr'][10]."""
    from oracle import harness, nmma_oracle as O
    inj = json.load(open(os.path.join(GOLDEN, "bu2019lm_injection.json")))
    row = {k: v[0] for k, v in inj.items()}
    sample_times = np.arange(0.1, 10.0 + 0.5, 0.5)
    model = O.OracleSVDLightCurveModel(["log10_mej_dyn", "log10_mej_wind", "KNtheta"], harness.oracle_core(core),
                                       ["ztfr"], sample_times)
    params = model.parameter_conversion(dict(row))
    tobs, lc = model.gen_detector_lc(params)
    trigger = 44244.0 + row["geocent_time"] / 86400.0 + row["timeshift"]   # GPS epoch, no leap seconds in 1980
    keep = tobs >= 0
    mags = lc["ztfr"][keep]
    noise = np.random.default_rng(42).normal(scale=0.1, size=len(mags))
    return (tobs[keep] + trigger)[10], (mags + noise)[10], model.redshift


def test_golden_value_from_committed_fixture():
    t10, m10, z = _golden_mag(fixture_core("mlp", ("ztfr",)))
    assert np.isclose([t10, m10, 0.1], [4.4248125e04, 2.09294036584e01, 1.0e-01]).all()   # the reference's assertion
    assert abs(m10 - 20.9294036584) < 5e-6                                               # and much tighter
    assert abs(z - 0.0115438) < 1e-6


@pytest.mark.reference
def test_golden_value_from_reference_files():
    """Same, but reading the reference's own .joblib / .h5 files with the product loaders."""
    from nmma_b200.mlmodel import load_surrogate
    core, filters, found, kind = load_surrogate("Bu2019nsbh", f"{REFERENCE}/nmma/tests/data", filters=["ztfr"],
                                                interpolation_type="tensorflow")
    assert kind == "mlp" and found == ["ztfr"]
    _, m10, _ = _golden_mag(core)
    assert abs(m10 - 20.9294036584) < 5e-6


@pytest.mark.reference
def test_fixture_npz_matches_reference_files():
    from nmma_b200.mlmodel import load_keras_mlp, load_sklearn_gps, load_svd_core
    core = load_svd_core(f"{REFERENCE}/nmma/tests/data/Bu2019nsbh.joblib")
    assert len(core) == 26 and "ps1::g" in core and "uvot::white" in core
    fx = fixture_core("mlp")
    for f in ("ztfr", "sdssu", "2massks"):
        W1, b1, W2, b2 = load_keras_mlp(f"{REFERENCE}/nmma/tests/data/Bu2019nsbh_tf/{f}.h5")
        assert W1.shape == (3, 2048) and W2.shape == (2048, 10) and W1.dtype == np.float32
        for a, b in zip((W1, b1, W2, b2), fx[f]["model"]):
            assert np.array_equal(a, b)
        assert np.array_equal(core[f]["VA"][:, :10], fx[f]["VA"][:, :10])
    gp = load_sklearn_gps(f"{REFERENCE}/nmma/tests/data/Bu2019nsbh/ztfr.joblib")
    assert gp["X"].shape == (891, 3) and gp["alpha"].shape == (10, 891)
    assert np.allclose(np.sqrt(gp["c2"][:3]), [0.535, 2.87, 2.68], rtol=2e-3)


@pytest.mark.reference
def test_unpacked_gp_equals_sklearn_predict():
    """kernel_(x, X_train_) @ alpha_ rebuilt from the unpacked arrays == GaussianProcessRegressor.predict."""
    import warnings
    import joblib
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gps = joblib.load(f"{REFERENCE}/nmma/tests/data/Bu2019nsbh/ztfr.joblib")
    ogps = oracle_ready_core(fixture_core("gp", ("ztfr",)))["ztfr"]["gps"]
    rng = np.random.default_rng(0)
    for x in rng.uniform(-0.2, 1.3, size=(12, 3)):
        for i in range(10):
            a = gps[i].predict(np.atleast_2d(x), return_std=True)[0].item()
            b = ogps[i].predict(np.atleast_2d(x)).item()
            assert abs(a - b) <= 1e-9 * max(1.0, abs(a))


def test_scipy_edge_semantics_used_by_the_oracle():
    """the SciPy behaviours the likelihood relies on (pinned against this SciPy)."""
    from scipy.stats import norm, truncnorm
    assert np.nan_to_num(-np.inf) == -1.7976931348623157e308
    assert truncnorm.logpdf(17.4, -np.inf, np.inf, loc=17.5, scale=1) == pytest.approx(norm.logpdf(17.4, 17.5, 1))
    with np.errstate(all="ignore"):
        assert np.isnan(truncnorm.logpdf(18, -np.inf, np.nan, loc=np.inf, scale=1))
        assert truncnorm.logpdf(23.0, -np.inf, (22 - 17.5) / 1.0, loc=17.5, scale=1.0) == -np.inf
        assert norm.logsf(19.6, np.inf, 1) == 0.0
    assert truncnorm.logpdf(17.4, -np.inf, 4.5, loc=17.5, scale=1) == pytest.approx(-0.923935135525776, rel=1e-12)
    assert norm.logsf(19.6, 18, 1) == pytest.approx(-2.904078010302249, rel=1e-12)
    assert norm.logsf(19.6, 0, 0.5) == pytest.approx(-772.9082649951413, rel=1e-12)


def test_oracle_autocomplete_matches_np_interp_semantics():
    from oracle.nmma_oracle import autocomplete_data
    xp = np.array([0.0, 1.0, 2.0, 3.0])
    fp = np.array([1.0, np.inf, 3.0, 5.0])
    out = autocomplete_data(np.array([-1.0, 0.5, 2.0, 2.5, 4.0]), xp, fp, extrapolate=np.inf)
    assert np.array_equal(out, [np.inf, 1.5, 3.0, 4.0, np.inf])          # the inf node is dropped, not propagated
    assert np.isinf(autocomplete_data(np.array([1.0]), xp, np.array([np.nan, np.nan, np.nan, 2.0]), np.inf)).all()
    out = autocomplete_data(np.array([-1.0, 5.0]), xp, np.array([1.0, 2.0, 3.0, 4.0]), extrapolate="constant")
    assert np.array_equal(out, [1.0, 4.0])


def test_oracle_sentinel_paths():
    """Detection outside the model window -> sentinel; upper limit outside -> contributes 0."""
    from oracle import harness
    from nmma_b200.core.priors import PriorDict, Uniform
    core = fixture_core("mlp", ("ztfr",))
    times = {"ztfr": np.array([1.0, 30.0])}
    mags = {"ztfr": np.array([20.0, 22.0])}
    priors = PriorDict()
    priors["luminosity_distance"] = Uniform(20.0, 60.0)
    priors["KNtheta"] = Uniform(0.0, 90.0)
    priors["log10_mej_dyn"] = Uniform(-2.0, -1.05)
    priors["log10_mej_wind"] = Uniform(-2.0, -1.05)
    cols = list(priors.keys())
    pts = np.array([[40.0, 30.0, -1.5, -1.5]])
    for err2, expect_sentinel in ((0.2, True), (np.inf, False)):
        errs = {"ztfr": np.array([0.1, err2])}
        lik, fixed = harness.build_oracle_likelihood(core, ["log10_mej_dyn", "log10_mej_wind", "KNtheta"], ["ztfr"],
                                                     core["ztfr"]["tt"], ["ztfr"], (times, mags, errs, 0.0), priors)
        val = harness.oracle_logl(lik, fixed, pts, cols)[0]
        assert (val == -1.7976931348623157e308) == expect_sentinel
