"""Derive the small committed fixtures from the reference checkout (run once, in the build container).

    python tests/golden/make_fixtures.py

Outputs (committed; /root/reference does not exist on the GPU box):
  nmma_b200/data/at2017gfo.json           AT2017gfo photometry parsed from example_files/lightcurves/AT2017gfo.dat
  tests/golden/bu2019nsbh_fixture.npz     Bu2019nsbh test surrogate: SVD basis (first 10 columns of VA), the three
                                          Keras MLPs (ztfr, sdssu, 2massks) and the ztfr scikit-learn GPs, unpacked
  tests/golden/bu2019lm_injection.json    the 100 Bu2019lm prior draws of nmma/tests/data/Bu2019lm_injection.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from nmma_b200.em.io import load_em_observations  # noqa: E402
from nmma_b200.mlmodel import load_keras_mlp, load_sklearn_gps, load_svd_core  # noqa: E402


def main():
    data = load_em_observations(f"{REF}/example_files/lightcurves/AT2017gfo.dat")
    out = {f: {k: [float(x) if np.isfinite(x) else "inf" for x in v] for k, v in d.items()} for f, d in data.items()}
    with open(f"{ROOT}/nmma_b200/data/at2017gfo.json", "w") as fh:
        json.dump({"source": "nmma example_files/lightcurves/AT2017gfo.dat (time in MJD, UTC)",
                   "trigger_time_mjd": 57982.5285236896, "data": out}, fh, indent=0)

    core = load_svd_core(f"{REF}/nmma/tests/data/Bu2019nsbh.joblib")
    arrays = {}
    mlp_filters = ["ztfr", "sdssu", "2massks"]
    for f in mlp_filters:
        c = core[f]
        key = f.replace(":", "_")
        arrays[f"{key}/VA"] = np.ascontiguousarray(c["VA"][:, :10])
        for name in ("mins", "maxs", "tt", "param_mins", "param_maxs"):
            arrays[f"{key}/{name}"] = np.asarray(c[name], float)
        W1, b1, W2, b2 = load_keras_mlp(f"{REF}/nmma/tests/data/Bu2019nsbh_tf/{f}.h5")
        arrays[f"{key}/W1"], arrays[f"{key}/b1"], arrays[f"{key}/W2"], arrays[f"{key}/b2"] = W1, b1, W2, b2
    gp = load_sklearn_gps(f"{REF}/nmma/tests/data/Bu2019nsbh/ztfr.joblib")
    for k, v in gp.items():
        arrays[f"ztfr/gp_{k}"] = v
    arrays["filters"] = np.array(mlp_filters)
    np.savez_compressed(f"{ROOT}/tests/golden/bu2019nsbh_fixture.npz", **arrays)

    inj = json.load(open(f"{REF}/nmma/tests/data/Bu2019lm_injection.json"))["injections"]
    content = inj.get("content", inj)
    with open(f"{ROOT}/tests/golden/bu2019lm_injection.json", "w") as fh:
        json.dump({k: v for k, v in content.items()}, fh)
    print("fixtures written")


if __name__ == "__main__":
    main()
