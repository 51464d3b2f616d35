"""Shared builders for the parity tests: one in-memory configuration -> GPU likelihood + oracle."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

SENTINEL = -1.7976931348623157e308


def fixture_core(kind="mlp", filters=("ztfr", "sdssu", "2massks")):
    """Bu2019nsbh fixture weights (tests/golden/bu2019nsbh_fixture.npz) in the reference's in-memory layout."""
    z = np.load(os.path.join(GOLDEN, "bu2019nsbh_fixture.npz"))
    core = {}
    for f in filters:
        key = f.replace(":", "_")
        T = z[f"{key}/VA"].shape[0]
        VA = np.zeros((T, T))
        VA[:, :10] = z[f"{key}/VA"]
        entry = dict(VA=VA, mins=z[f"{key}/mins"], maxs=z[f"{key}/maxs"], tt=z[f"{key}/tt"],
                     param_mins=z[f"{key}/param_mins"], param_maxs=z[f"{key}/param_maxs"], n_coeff=10)
        if kind == "mlp":
            entry["model"] = tuple(z[f"{key}/{n}"] for n in ("W1", "b1", "W2", "b2"))
        else:
            entry["gps"] = {k: z[f"ztfr/gp_{k}"] for k in ("X", "alpha", "c2", "rq_alpha", "rq_len", "ymean", "ystd")}
        core[f] = entry
    return core


from oracle.harness import UnpackedGP, oracle_ready_core  # noqa: E402,F401  (shared with bench.py's CPU legs)


def synthetic_observations(filters, rng, n_per_filter=8, tmin=0.3, tmax=12.0, n_ul=1, mag0=18.0, slope=0.4):
    """(times{f}, mags{f}, errs{f}, trigger) relative to trigger, with `n_ul` upper limits per filter."""
    times, mags, errs = {}, {}, {}
    for f in filters:
        n = int(n_per_filter if np.isscalar(n_per_filter) else n_per_filter[f])
        t = np.sort(rng.uniform(tmin, tmax, n))
        m = mag0 + slope * t + rng.normal(scale=0.3, size=n)
        e = rng.uniform(0.02, 0.3, n)
        idx = rng.choice(n, size=min(n_ul, n), replace=False) if n_ul else []
        for i in idx:
            e[i] = np.inf
        times[f], mags[f], errs[f] = t, m, e
    return times, mags, errs, 0.0


def build_pair(core, model_name, model_filters, obs_filters, lc_data, priors, kind="mlp",
               sample_times=None, error_budget=1.0, systematics=None, detection_limit=np.inf,
               model_parameters=None, extinction_law=None, extinction_coef=None):
    """(GPU likelihood, oracle likelihood, fixed dict, columns)."""
    from nmma_b200.em import EMTransientLikelihood, FilterSystematicsHandler, SVDLightCurveModel
    from oracle import harness

    itype = "sklearn_gp" if kind == "gp" else "tensorflow"
    model = SVDLightCurveModel(model_name, svd_mag_model=core, interpolation_type=itype,
                               filters=list(model_filters), sample_times=sample_times,
                               model_parameters=model_parameters, extinction_law=extinction_law)
    if extinction_coef is not None:
        model.extinction_coefficients = dict(extinction_coef)
    handler = FilterSystematicsHandler(list(obs_filters), systematics, error_budget, lc_data[0])
    if systematics is not None:
        handler.setup_systematics_priors(priors)
    lik = EMTransientLikelihood(model, lc_data, handler, priors, filters=list(obs_filters),
                                detection_limit=detection_limit)
    plan = handler.device_plan()
    olik, fixed = harness.build_oracle_likelihood(
        oracle_ready_core(core), model.model_parameters, list(model_filters), np.asarray(model.model_times, float),
        list(obs_filters), lc_data, priors, sys_plan=plan, detection_limit=detection_limit,
        z_table=model._z_table, filts_lambdas=(model.default_filts, model.lambdas), extinction_law=extinction_law,
        extinction_coef=extinction_coef)
    return lik, olik, fixed, lik.columns


def assert_logl_close(got, ref, rtol=1e-4):
    """north-star tolerance: |dlogL| <= 1e-4 * max(1, |logL|); sentinel / failure masks bit-exact."""
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    sg, sr = got == SENTINEL, ref == SENTINEL
    assert np.array_equal(sg, sr), f"sentinel masks differ at {np.nonzero(sg != sr)[0][:10]}"
    ok = ~sr
    err = np.abs(got[ok] - ref[ok]) / np.maximum(1.0, np.abs(ref[ok]))
    assert err.size == 0 or err.max() <= rtol, f"max rel err {err.max():.3e} (> {rtol})"
    return float(err.max()) if err.size else 0.0


# ---------------------------------------------------------------------------------------------
# Golden vectors produced by the reference's own source files (tests/golden/make_reference_vectors.py)
# ---------------------------------------------------------------------------------------------
_YAML_WITH_TIME = {"config": {"withTime": {"value": True, "filters": [None], "time_nodes": 4, "type": "Uniform",
                                            "minimum": 0, "maximum": 2},
                              "withoutTime": {"value": False, "type": "Uniform", "minimum": 0, "maximum": 2}}}
_YAML_WITHOUT_TIME = {"config": {"withTime": {"value": False, "filters": [None], "time_nodes": 4, "type": "Uniform",
                                               "minimum": 0, "maximum": 2},
                                 "withoutTime": {"value": True, "type": "Uniform", "minimum": 0, "maximum": 2}}}
REFERENCE_CASES = {
    "A": dict(kind="mlp", filters=["ztfr", "sdssu", "2massks"], sample_times=None, budget=1.0, yaml=None, limit=np.inf),
    "B": dict(kind="mlp", filters=["ztfr", "sdssu", "2massks"], sample_times=np.arange(0.1, 10.0 + 0.5, 0.5),
              budget=0.5, yaml=None, limit=np.inf),
    "C": dict(kind="mlp", filters=["ztfr", "sdssu", "2massks"], sample_times=None, budget=1.0, yaml=_YAML_WITH_TIME,
              limit={"ztfr": 23.5, "sdssu": 24.0, "2massks": 23.0}),
    "D": dict(kind="gp", filters=["ztfr"], sample_times=None, budget=0.8, yaml=None, limit=np.inf),
    "E": dict(kind="mlp", filters=["ztfr", "sdssu", "2massks"], sample_times=None, budget=1.0, yaml=_YAML_WITHOUT_TIME,
              limit=np.inf),
}


def reference_case(name):
    """Rebuild case `name` of make_reference_vectors.py: (cfg, lc_data, priors, golden dict)."""
    import copy
    from nmma_b200.core.priors import PriorDict, Sine, Uniform
    z = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    cfg = copy.deepcopy(REFERENCE_CASES[name])
    times, mags, errs = {}, {}, {}
    for f in cfg["filters"]:
        key = f.replace(":", "_")
        times[f], mags[f], errs[f] = (z[f"{name}/obs/{key}/{k}"] for k in ("time", "mag", "mag_error"))
    priors = PriorDict()
    priors["luminosity_distance"] = Uniform(10.0, 200.0, name="luminosity_distance")
    priors["inclination_EM"] = Sine(0.0, np.pi / 2, name="inclination_EM")
    priors["timeshift"] = Uniform(-0.3, 0.3, name="timeshift")
    priors["log10_mej_dyn"] = Uniform(-2.2, -0.9, name="log10_mej_dyn")
    priors["log10_mej_wind"] = Uniform(-2.2, -0.9, name="log10_mej_wind")
    gold = {k: z[f"{name}/{k}"] for k in ("points", "columns", "logl", "mags", "tobs", "z_table")}
    gold["columns"] = [str(c) for c in gold["columns"]]
    return cfg, (times, mags, errs, 57000.0), priors, gold


def build_reference_pair(name):
    """(GPU likelihood, oracle likelihood, fixed, cols, golden) for a reference-generated case."""
    cfg, lc_data, priors, gold = reference_case(name)
    core = fixture_core(cfg["kind"], tuple(cfg["filters"]))
    lik, olik, fixed, cols = build_pair(core, "Bu2019nsbh", cfg["filters"], cfg["filters"], lc_data, priors,
                                        kind=cfg["kind"], sample_times=cfg["sample_times"], error_budget=cfg["budget"],
                                        systematics=cfg["yaml"], detection_limit=cfg["limit"])
    assert cols == gold["columns"], (cols, gold["columns"])
    # evaluate with the very dL -> z table the reference run used
    table = (gold["z_table"][0], gold["z_table"][1])
    lik.sub_model.light_curve_model._z_table = table
    lik.sub_model.light_curve_model.redshift_func = lambda p: np.interp(p["luminosity_distance"], *table)
    lik.sub_model._engine = None
    olik.light_curve_model.check_vs_priors(table=table)
    return lik, olik, fixed, cols, gold
