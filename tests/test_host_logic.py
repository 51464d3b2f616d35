"""Host-side mirror of the reference interface: priors, systematics naming contract, io, cosmology, layout."""
import copy
import os
import pickle

import numpy as np
import pytest

from conftest import REFERENCE
from helpers import fixture_core, synthetic_observations


# ---- systematics: the naming contract pinned by nmma/tests/systematics.py ---------------------
def test_systematics_prior_strings_match_reference_contract():
    from nmma_b200.em import systematics as s
    vals = {"type": "Uniform", "minimum": 0.0, "maximum": 1.0, "time_nodes": 2, "filters": [["bessellb", "bessellv"], "ztfr"]}
    res = s.handle_withTime(vals)
    assert "em_syserr_bessellb___bessellv_0" in res[0] and "em_syserr_ztfr_1" in res[3] and len(res) == 4
    res = s.handle_withoutTime({"type": "Uniform", "minimum": 0.0, "maximum": 1.0})
    assert res == ["em_syserr = Uniform(minimum=0.0, maximum=1.0, name='em_syserr', latex_label='em_syserr', unit=None, boundary=None)"]
    res = s.handle_withTime({"type": "Uniform", "minimum": 0.0, "maximum": 1.0, "time_nodes": 1, "filters": [None]})
    assert len(res) == 1 and "em_syserr_all_0" in res[0]
    for f in s.ALLOWED_FILTERS:
        r = s.handle_withTime({"type": "Uniform", "minimum": 0.0, "maximum": 1.0, "time_nodes": 1, "filters": [f]})
        assert len(r) == 1 and f"em_syserr_{f}_0" in r[0]


def test_systematics_validation_errors():
    from nmma_b200.em import systematics as s
    with pytest.raises(s.ValidationError, match="Only one configuration key can be set to True at a time"):
        s.validate_only_one_true({"config": {"withTime": {"value": True}, "withoutTime": {"value": True}}})
    with pytest.raises(s.ValidationError, match="At least one configuration key must be set to True"):
        s.validate_only_one_true({"config": {"withTime": {"value": False}, "withoutTime": {"value": False}}})
    with pytest.raises(s.ValidationError, match="'value' key must be present and be a boolean"):
        s.validate_only_one_true({"config": {"withTime": {}, "withoutTime": {"value": False}}})
    with pytest.raises(s.ValidationError, match="Invalid filter value 'invalid_filter'"):
        s.validate_filters([["bessellb", "invalid_filter"], "ztfr"])
    with pytest.raises(s.ValidationError, match="Duplicate filter value 'bessellb' within the same group"):
        s.validate_filters([["bessellb", "bessellb"], "ztfr"])
    with pytest.raises(s.ValidationError, match="Duplicate filter value 'bessellb'. A filter can only be used in one group"):
        s.validate_filters([["bessellb", "bessellv"], "bessellb"])
    s.validate_filters([["bessellb", "bessellv"], None])
    s.validate_filters([])
    with pytest.raises(KeyError):
        s.ALLOWED_DISTRIBUTIONS["uniform"]
    with pytest.raises(FileNotFoundError):
        s.main("non_existent_file.yaml")


@pytest.mark.reference
def test_shipped_systematics_yaml_files():
    from nmma_b200.em import systematics as s
    res = s.main(f"{REFERENCE}/priors/systematics.yaml")           # ships min/max instead of minimum/maximum
    assert len(res) == 8 and res[0].startswith("em_syserr_sdssu_0 = Uniform(minimum=0, maximum=2")
    assert any("em_syserr_2massj___2massh_1" in r for r in res) and any("em_syserr_all_0" in r for r in res)
    res = s.main(f"{REFERENCE}/nmma/tests/data/systematics_with_time_combined_filters.yaml")
    assert len(res) == 12
    assert s.main(f"{REFERENCE}/nmma/tests/data/systematics_without_time.yaml")[0].startswith("em_syserr = Uniform")


def test_filter_systematics_handler_plans():
    from nmma_b200.core.priors import PriorDict
    from nmma_b200.em.systematics import FilterSystematicsHandler
    filters = ["sdssu", "2massj", "2massh", "2massks", "ps1::g"]
    times = {f: np.linspace(0.5, 10, 4) for f in filters}
    tt = np.arange(0, 21.01, 0.1)
    # constant budget (float, list, dict, string)
    for budget, expect in ((None, 1.0), (0.3, 0.3), ("0.25", 0.25), ([0.1, 0.2, 0.3, 0.4, 0.5], 0.1), ({"sdssu": 0.7}, 0.7)):
        h = FilterSystematicsHandler(filters, None, budget, times)
        h.reset(tt, PriorDict())
        assert h.device_plan()["sdssu"] == ("budget", expect)
    # sampled em_syserr without YAML
    p = PriorDict({"em_syserr": "Uniform(minimum=0, maximum=2)"})
    h = FilterSystematicsHandler(filters, None, 1.0, times)
    h.reset(tt, p)
    assert h.device_plan()["ps1::g"] == ("param", "em_syserr")
    # legacy YAML: the null entry sends every filter to 'all' and stops (reference quirk)
    yml = {"config": {"withTime": {"value": True, "filters": ["sdssu", None, ["2massj", "2massh"]], "time_nodes": 4,
                                   "type": "Uniform", "minimum": 0, "maximum": 2},
                      "withoutTime": {"value": False, "type": "Uniform", "minimum": 0, "maximum": 2}}}
    p = PriorDict()
    h = FilterSystematicsHandler(filters, copy.deepcopy(yml), 1.0, times)
    h.setup_systematics_priors(p)
    assert "em_syserr_2massj___2massh_3" in p and "em_syserr_all_0" in p          # sampled-but-unused priors exist
    h.reset(tt, p)
    plan = h.device_plan()
    assert all(plan[f][1] == [f"em_syserr_all_{i}" for i in range(4)] for f in filters)
    assert np.array_equal(plan["sdssu"][2], [0.0, 7.0, 14.0, 21.0])
    # host mirror of __call__ agrees with the plan
    params = {f"em_syserr_all_{i}": v for i, v in enumerate([0.1, 0.5, 0.9, 1.3])}
    sig = h(params)["sdssu"]
    assert np.allclose(sig, np.interp(times["sdssu"], [0, 7, 14, 21], [0.1, 0.5, 0.9, 1.3]))
    # new-style YAML with a missing prior -> assertion like the reference (systematics.py:271,276)
    h2 = FilterSystematicsHandler(filters, {"time_nodes": 3, "prior": "Uniform(minimum=0, maximum=1)"}, 1.0, times)
    with pytest.raises(AssertionError):
        h2.reset(tt, PriorDict())
    p3 = PriorDict()
    h2.setup_systematics_priors(p3)
    assert list(p3) == ["em_syserr_0", "em_syserr_1", "em_syserr_2"]
    h2.reset(tt, p3)
    assert np.allclose(h2.device_plan()["2massks"][2], [0.0, 10.5, 21.0])


# ---- priors -------------------------------------------------------------------------------------
@pytest.mark.reference
def test_prior_files_parse():
    from nmma_b200.core.priors import PriorDict
    p = PriorDict(f"{REFERENCE}/priors/Bu2019lm.prior")
    assert list(p.keys()) == ["luminosity_distance", "KNphi", "inclination_EM", "timeshift", "log10_mej_dyn", "log10_mej_wind"]
    assert p["timeshift"].minimum == -2.0 and p["inclination_EM"].maximum == pytest.approx(np.pi / 2)
    p = PriorDict(f"{REFERENCE}/priors/Bu2023Ye.prior")
    assert len(p) == 9 and p["Yewind"].maximum == 0.40
    p = PriorDict(f"{REFERENCE}/example_files/prior/ZTF_Bu2019lm.prior")
    assert p.fixed_keys, "bare float literals are fixed values"


def test_prior_sampling_and_rescale():
    from nmma_b200.core.priors import Constraint, DeltaFunction, Interped, PriorDict, Sine, Uniform
    p = PriorDict({"a": "Uniform(minimum=-1, maximum=3)", "b": "Sine(minimum=0, maximum=np.pi/2)", "c": "0.5",
                   "d": "LogUniform(minimum=1e-3, maximum=1)", "e": "Gaussian(mu=1, sigma=2)"})
    assert p.non_fixed_keys == ["a", "b", "d", "e"] and p.fixed_keys == ["c"]
    pts, keys = p.sample_array(20000, np.random.default_rng(0))
    assert keys == ["a", "b", "d", "e"] and pts.shape == (20000, 4)
    assert -1 <= pts[:, 0].min() and pts[:, 0].max() <= 3 and abs(pts[:, 0].mean() - 1) < 0.05
    assert abs(np.cos(pts[:, 1]).mean() - 0.5) < 0.01                     # p(x) ~ sin x on [0, pi/2]
    assert abs(np.log10(pts[:, 2]).mean() + 1.5) < 0.03
    assert abs(pts[:, 3].std() - 2) < 0.05
    assert Sine(0, np.pi).rescale(0.5) == pytest.approx(np.pi / 2)
    assert DeltaFunction(2.0).rescale(0.3) == 2.0
    tri = Interped([0, 0.5], [4.0, 0.0])                                  # the reference's Ebv prior shape
    assert 0 <= tri.rescale(0.5) <= 0.5 and tri.prob(0.6) == 0
    c = Constraint(0, 1)
    assert c.prob(0.5) and not c.prob(1.5)
    with pytest.raises(ValueError):
        PriorDict({"x": "bilby.gw.prior.AlignedSpin(name='chi_1')"})
    assert pickle.loads(pickle.dumps(p)).keys() == p.keys()


# ---- io -------------------------------------------------------------------------------------------
@pytest.mark.reference
def test_at2017gfo_reader_and_fixture():
    from nmma_b200 import synthetic as syn
    from nmma_b200.em.io import load_em_observations
    d = load_em_observations(f"{REFERENCE}/example_files/lightcurves/AT2017gfo.dat")
    assert sum(len(v["time"]) for v in d.values()) == 141
    assert {f: len(v["time"]) for f, v in d.items()} == {"ps1::g": 13, "ps1::r": 19, "ps1::i": 20, "ps1::z": 18, "ps1::y": 15,
                                                         "2massj": 14, "2massh": 17, "2massks": 23, "sdssu": 2}
    assert sum(int(np.isinf(v["mag_error"]).sum()) for v in d.values()) == 3
    lc, filters = syn.load_at2017gfo(data_tmax=np.inf)
    for f in filters:
        assert np.array_equal(lc[0][f], d[f]["time"] - syn.AT2017GFO_TRIGGER_MJD)
        assert np.array_equal(lc[1][f], d[f]["mag"])


def test_time_parsing(tmp_path):
    from nmma_b200.em.io import gps_to_mjd, isot_to_mjd, load_em_observations
    assert isot_to_mjd("2017-08-17T12:41:04.4") == pytest.approx(57982.52852314815, abs=1e-9)
    assert gps_to_mjd(1187008882.4) == pytest.approx(57982.52852314815, abs=1e-9)
    assert gps_to_mjd(0.0) == 44244.0
    f = tmp_path / "lc.dat"
    f.write_text("# c\n57983.0 ztfg 17.4 0.02\n2017-08-18T12:00:00 ztfr 18.0 inf\n\n57984.5 ztfg 18.1 0.05\n")
    d = load_em_observations(str(f))
    assert list(d) == ["ztfg", "ztfr"] and d["ztfr"]["time"][0] == 57983.5 and np.isinf(d["ztfr"]["mag_error"][0])
    assert np.array_equal(d["ztfg"]["time"], [57983.0, 57984.5])


# ---- cosmology ---------------------------------------------------------------------------------------
def test_cosmology_product_vs_oracle():
    """Two independent restatements of astropy's Planck18 (product: Gauss-Legendre, oracle: quad + brentq)."""
    from nmma_b200.core.cosmology import Planck18 as P
    from nmma_b200.core.conversion import get_cosmo_grids
    from oracle.cosmology import Planck18 as O, get_cosmo_grids as ogrids
    for z in (1e-4, 0.0115, 0.05, 0.5, 2.0):
        assert P.luminosity_distance(z) == pytest.approx(O.luminosity_distance(z), rel=1e-11)
    assert P.z_at_luminosity_distance(51.600230897327414) == pytest.approx(0.011543898839, rel=1e-9)
    dg, zg = get_cosmo_grids(1.0, 200.0)
    odg, ozg = ogrids(1.0, 200.0)
    assert len(zg) == 50 and np.allclose(zg, ozg, rtol=1e-10) and np.allclose(dg, odg, rtol=1e-10)
    assert P.Ode0 == pytest.approx(0.6888463055445441, rel=1e-6)          # astropy's Planck18.Ode0
    with pytest.raises(ValueError):
        get_cosmo_grids(0.0, 200.0)                                       # geomspace(0, ...) in the reference


# ---- parameter layout -----------------------------------------------------------------------------------
def test_resolve_param_sources():
    from nmma_b200 import _lib as L
    from nmma_b200._lib import ParamSrc
    from nmma_b200.em.model import model_parameters_dict, resolve_param_sources
    avail = {k: ParamSrc.column(i) for i, k in enumerate(["luminosity_distance", "KNphi", "inclination_EM", "timeshift",
                                                          "log10_mej_dyn", "log10_mej_wind"])}
    src = resolve_param_sources(model_parameters_dict["Bu2019lm"], avail)
    assert [(s.col, s.transform) for s in src] == [(4, 0), (5, 0), (1, 0), (2, L.XF_RAD2DEG)]
    src = resolve_param_sources(["log10_mej", "vej"], {"mej": ParamSrc.column(0), "log10_vej": ParamSrc.const(-1.0)})
    assert (src[0].col, src[0].transform) == (0, L.XF_LOG10) and (src[1].col, src[1].transform, src[1].value) == (-1, L.XF_POW10, -1.0)
    src = resolve_param_sources(["KNtheta"], {"cos_theta_jn": ParamSrc.column(2)})
    assert src[0].transform == L.XF_COSTHETAJN_DEG
    src = resolve_param_sources(["KNtheta"], {})
    assert src[0].col == -1 and src[0].value == 0.0
    with pytest.raises(AttributeError):
        resolve_param_sources(["Yewind"], avail)


def test_filter_name_mapping():
    from nmma_b200.em.utils import average_mags, get_filter_name_mapping
    direct, avg = get_filter_name_mapping(["ps1::g", "B", "w", "V", "radio-3GHz", "UVW2"])
    assert direct == {"ps1::g": "ps1::g", "B": "g", "radio-3GHz": "radio-3GHz", "UVW2": "u"}
    assert avg == {"w": ["g", "r", "i"], "V": ["g", "r"]}
    with pytest.raises(ValueError, match="Unknown filter"):
        get_filter_name_mapping(["not_a_filter"])
    assert average_mags({"g": 1.0, "r": 2.0, "i": 6.0}, "w") == 3.0


def test_model_and_likelihood_construct_and_pickle_without_gpu():
    """Object construction, layouts and pickling are host-only; device handles never enter a pickle."""
    from nmma_b200.core.priors import PriorDict, Sine, Uniform
    from nmma_b200.em import EMTransientLikelihood, FilterSystematicsHandler, OpticalLightCurve, SVDLightCurveModel
    filters = ["ztfr", "sdssu", "2massks"]
    core = fixture_core("mlp", filters)
    model = SVDLightCurveModel("Bu2019nsbh", svd_mag_model=core, interpolation_type="tensorflow", filters=filters)
    assert model.model_parameters == ["log10_mej_dyn", "log10_mej_wind", "KNtheta"]
    assert len(model.model_times) == 211 and model.weights.W1.shape == (3, 3, 2048)
    lc = synthetic_observations(filters, np.random.default_rng(0))
    priors = PriorDict()
    priors["luminosity_distance"] = Uniform(10.0, 200.0)
    priors["inclination_EM"] = Sine(0.0, np.pi / 2)
    priors["log10_mej_dyn"] = Uniform(-2.0, -1.05)
    priors["log10_mej_wind"] = 1.5 - 3.0                                   # fixed value
    handler = FilterSystematicsHandler(filters, None, 0.5, lc[0])
    lik = EMTransientLikelihood(model, lc, handler, priors, filters=filters, detection_limit={"ztfr": 22.0})
    assert lik.columns == ["luminosity_distance", "inclination_EM", "log10_mej_dyn"]
    assert lik.noise_log_likelihood() == 0.0 and model._z_table is not None
    assert lik.sub_model.detection_limit == {"ztfr": 22.0, "sdssu": np.inf, "2massks": np.inf}
    plan = lik.sub_model.plan_layout(lik.columns)                          # host-only layout resolution
    assert plan["zmode"] == 2 and plan["P"] == 3 and [s.col for s in plan["xsrc"]] == [2, -1, 1]
    assert plan["xsrc"][1].value == -1.5 and plan["obs"][0] == [[0], [1], [2]] and plan["sys"][1] == [0.5] * 3
    lik2 = pickle.loads(pickle.dumps(lik))
    assert lik2.sub_model._engine is None and lik2.columns == lik.columns
    p = lik.parameter_conversion({"inclination_EM": np.pi / 4, "log10_mej_dyn": -1.5})
    assert p["KNtheta"] == pytest.approx(45.0)
    with pytest.raises(ValueError, match="interpolation-type"):
        SVDLightCurveModel("Bu2019nsbh", svd_mag_model=core, interpolation_type="nonsense", filters=filters)
    with pytest.raises(ValueError, match="Multiple equivalent parameters"):
        bad = PriorDict({"inclination_EM": "Uniform(minimum=0, maximum=1)", "KNtheta": "Uniform(minimum=0, maximum=90)"})
        EMTransientLikelihood(model, lc, FilterSystematicsHandler(filters, None, 0.5, lc[0]), bad, filters=filters)
    old = OpticalLightCurve(model, filters, {f: np.c_[lc[0][f] + 100.0, lc[1][f], lc[2][f]] for f in filters}, 100.0,
                            error_budget=1.0, tmin=0.0, tmax=14.0, priors=priors)
    assert old.columns == lik.columns


@pytest.mark.reference
def test_model_loads_reference_layout_from_disk():
    from nmma_b200.em import SVDLightCurveModel
    m = SVDLightCurveModel("Bu2019nsbh", svd_path=f"{REFERENCE}/nmma/tests/data", interpolation_type="tensorflow",
                           filters=["ztfr", "sdssu"], local_only=True)
    assert m._eval_filters == ["ztfr", "sdssu"] and m.weights.kind == "mlp"
    g = SVDLightCurveModel("Bu2019nsbh", svd_path=f"{REFERENCE}/nmma/tests/data", interpolation_type="sklearn_gp",
                           filters=["ztfr"], local_only=True)
    assert g.weights.kind == "gp" and g.weights.alpha.shape == (1, 10, 891)
    with pytest.raises(ValueError):
        SVDLightCurveModel("Bu2019nsbh", svd_path="/nonexistent", interpolation_type="tensorflow", filters=["ztfr"])
