"""Golden vectors produced by the reference's own code (run here, once; outputs committed).

    python tests/golden/make_reference_vectors.py

Executes ``nmma.em.model.SVDLightCurveModel``, ``nmma.em.systematics.FilterSystematicsHandler`` and
``nmma.em.em_likelihood.EMTransientLikelihood`` UNMODIFIED from /root/reference (see reference_stubs.py for
the two third-party stand-ins: Keras forward pass, astropy Planck18) on the Bu2019nsbh test surrogate shipped
with the reference, and stores inputs + outputs in tests/golden/reference_vectors.npz:

  case A  tensorflow MLP, 3 filters, constant error budget 1.0, default grid (tt), no detection limit
  case B  tensorflow MLP, --em-tmin 0.1 --em-tmax 10 --em-tstep 0.5 grid (two interpolation stages), budget 0.5
  case C  tensorflow MLP, legacy systematics YAML (withTime, 4 nodes, all filters) + finite detection limits
  case D  sklearn_gp (real GaussianProcessRegressor.predict), filter ztfr
  case E  legacy YAML withoutTime (one sampled em_syserr for all filters)

Each case: 96 prior draws (a few pushed out of the model's time window on purpose) -> reference
``log_likelihood(dict)``; plus ``gen_detector_lc`` magnitudes for the first 8 draws.
"""
import copy
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import reference_stubs as RS  # noqa: E402

REF = RS.REF
DATA = f"{REF}/nmma/tests/data"
YAML_WITH_TIME = {"config": {"withTime": {"value": True, "filters": [None], "time_nodes": 4, "type": "Uniform",
                                           "minimum": 0, "maximum": 2},
                             "withoutTime": {"value": False, "type": "Uniform", "minimum": 0, "maximum": 2}}}
YAML_WITHOUT_TIME = {"config": {"withTime": {"value": False, "filters": [None], "time_nodes": 4, "type": "Uniform",
                                              "minimum": 0, "maximum": 2},
                                "withoutTime": {"value": True, "type": "Uniform", "minimum": 0, "maximum": 2}}}


def observations(filters, seed):
    rng = np.random.default_rng(seed)
    data = {}
    for f in filters:
        n = 11
        t = np.sort(rng.uniform(0.4, 13.0, n))
        m = 18.5 + 0.25 * t + rng.normal(scale=0.3, size=n)
        e = rng.uniform(0.02, 0.3, n)
        e[rng.choice(n, size=2, replace=False)] = np.inf
        data[f] = {"time": t + 57000.0, "mag": m, "mag_error": e}
    return data


def base_priors(P):
    p = P.PriorDict()
    p["luminosity_distance"] = P.Uniform(10.0, 200.0, name="luminosity_distance")
    p["inclination_EM"] = P.Sine(0.0, np.pi / 2, name="inclination_EM")
    p["timeshift"] = P.Uniform(-0.3, 0.3, name="timeshift")
    p["log10_mej_dyn"] = P.Uniform(-2.2, -0.9, name="log10_mej_dyn")
    p["log10_mej_wind"] = P.Uniform(-2.2, -0.9, name="log10_mej_wind")
    return p


def main():
    mods = RS.load_reference()
    from nmma_b200.core import priors as P
    model_mod, lik_mod, sys_mod, utils_mod = mods["model"], mods["em_likelihood"], mods["systematics"], mods["utils"]

    cases = {
        "A": dict(itype="tensorflow", filters=["ztfr", "sdssu", "2massks"], sample_times=None, budget=1.0,
                  yaml=None, limit=np.inf),
        "B": dict(itype="tensorflow", filters=["ztfr", "sdssu", "2massks"], sample_times=np.arange(0.1, 10.0 + 0.5, 0.5),
                  budget=0.5, yaml=None, limit=np.inf),
        "C": dict(itype="tensorflow", filters=["ztfr", "sdssu", "2massks"], sample_times=None, budget=1.0,
                  yaml=YAML_WITH_TIME, limit={"ztfr": 23.5, "sdssu": 24.0, "2massks": 23.0}),
        "D": dict(itype="sklearn_gp", filters=["ztfr"], sample_times=None, budget=0.8, yaml=None, limit=np.inf),
        "E": dict(itype="tensorflow", filters=["ztfr", "sdssu", "2massks"], sample_times=None, budget=1.0,
                  yaml=YAML_WITHOUT_TIME, limit=np.inf),
    }
    out = {}
    for name, cfg in cases.items():
        filters = cfg["filters"]
        model = model_mod.SVDLightCurveModel("Bu2019nsbh", svd_path=DATA, interpolation_type=cfg["itype"],
                                             filters=list(filters), sample_times=cfg["sample_times"], local_only=True)
        raw = observations(filters, seed=ord(name))
        if name == "B":       # keep the data inside the shorter grid
            for f in raw:
                keep = raw[f]["time"] - 57000.0 < 8.5
                raw[f] = {k: v[keep] for k, v in raw[f].items()}
        lc_data = utils_mod.setup_filtered_lc_data(copy.deepcopy(raw), 57000.0)
        priors = base_priors(P)
        handler = sys_mod.FilterSystematicsHandler(list(filters), copy.deepcopy(cfg["yaml"]), cfg["budget"], lc_data[0])
        if cfg["yaml"] is not None:
            handler.setup_systematics_priors(priors)
        lik = lik_mod.EMTransientLikelihood(model, lc_data, handler, priors, filters=list(filters),
                                            detection_limit=cfg["limit"])
        cols = [k for k in priors.keys()]
        pts, _ = priors.sample_array(96, np.random.default_rng(1000 + ord(name)), cols)
        pts[90:, cols.index("timeshift")] = np.linspace(0.5, 14.0, 6)     # push detections out of the model window
        logl = np.array([lik.log_likelihood(dict(zip(cols, map(float, row)))) for row in pts])
        mags, tobs = [], []
        for row in pts[:8]:
            p = model.parameter_conversion(dict(zip(cols, map(float, row))))
            t, lc = model.gen_detector_lc(p)
            tobs.append(np.asarray(t, float))
            mags.append(np.stack([np.asarray(lc[f], float) for f in filters]))
        dist_grid, z_grid = mods["conversion"].get_cosmo_grids(10.0, 200.0, mods["constants"].get_cosmology())
        out[f"{name}/points"] = pts
        out[f"{name}/columns"] = np.array(cols)
        out[f"{name}/logl"] = logl
        out[f"{name}/mags"] = np.stack(mags)
        out[f"{name}/tobs"] = np.stack(tobs)
        out[f"{name}/z_table"] = np.stack([np.asarray(dist_grid, float), np.asarray(z_grid, float)])
        for f in filters:
            key = f.replace(":", "_")
            out[f"{name}/obs/{key}/time"] = lc_data[0][f]
            out[f"{name}/obs/{key}/mag"] = lc_data[1][f]
            out[f"{name}/obs/{key}/mag_error"] = lc_data[2][f]
        n_sent = int((logl == np.nan_to_num(-np.inf)).sum())
        print(f"case {name}: {len(cols)} columns, logL range [{logl[logl > -1e300].min():.3f}, {logl.max():.3f}], "
              f"{n_sent} sentinels")
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("written", os.path.join(HERE, "reference_vectors.npz"))


if __name__ == "__main__":
    main()
