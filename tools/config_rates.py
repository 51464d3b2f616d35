#!/usr/bin/env python
"""Device throughput of the other BASELINE.json configurations (not bench lines; parity for the same setups is in
tests/test_gpu_parity.py): C3 = Bu2023Ye-shaped MLP (d = 7) with time-node systematics and upper limits, with and without a
finite detection limit; C4 = Ka2017-shaped sklearn GP (Ntr = 329) across ZTF + PS1 filters with detection limits.
    python tools/config_rates.py [N_mlp] [N_gp]"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nmma_b200 import synthetic as syn
from nmma_b200.em import EMTransientLikelihood, FilterSystematicsHandler, SVDLightCurveModel

N_MLP = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
N_GP = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
YAML_TIME = {"config": {"withTime": {"value": True, "filters": [None], "time_nodes": 4, "type": "Uniform", "minimum": 0,
                                     "maximum": 2},
                        "withoutTime": {"value": False, "type": "Uniform", "minimum": 0, "maximum": 2}}}


def build(core, name, filters, lc_data, priors, kind="mlp", systematics=None, limit=np.inf):
    itype = "sklearn_gp" if kind == "gp" else "tensorflow"
    model = SVDLightCurveModel(name, svd_mag_model=core, interpolation_type=itype, filters=list(filters))
    handler = FilterSystematicsHandler(list(filters), systematics, 1.0, lc_data[0])
    if systematics is not None:
        handler.setup_systematics_priors(priors)
    return EMTransientLikelihood(model, lc_data, handler, priors, filters=list(filters), detection_limit=limit)


def rate(tag, lik, priors, n, reps=5):
    cols = lik.columns
    pts, _ = priors.sample_array(n, np.random.default_rng(5), cols)
    dev = torch.from_numpy(pts).cuda()
    eng = lik.sub_model.engine_for(cols)
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        eng.logl_device(dev, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.logl_device(dev, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sent = int((out.cpu().numpy() < -1e300).sum())
    print(f"{tag}: {ms:.3f} ms per {n} evals = {n / ms / 1e3:.2f} M evals/s (P = {len(cols)}, sentinel rows {sent})", flush=True)


lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
for lim in (np.inf, 24.5):
    core = syn.random_model("Bu2023Ye", filters, seed=1)
    priors = syn.bu2023ye_prior(); priors["timeshift"].maximum = 0.1
    lik = build(core, "Bu2023Ye", filters, lc_data, priors, systematics=copy.deepcopy(YAML_TIME), limit=lim)
    rate(f"C3 Bu2023Ye d=7, 4 time-node systematics, detection limit {lim}", lik, priors, N_MLP)
core = syn.random_model("Bu2023Ye", filters, seed=1)
priors = syn.bu2023ye_prior(); priors["timeshift"].maximum = 0.1
rate("C3' Bu2023Ye d=7, constant 1 mag budget", build(core, "Bu2023Ye", filters, lc_data, priors), priors, N_MLP)

gfilters = ["ztfg", "ztfr", "ztfi", "sdssu", "ps1::g", "ps1::r", "ps1::i", "ps1::z", "ps1::y"]
rng = np.random.default_rng(8)
times, mags, errs = {}, {}, {}
for f in gfilters:
    t = np.sort(rng.uniform(0.3, 13.0, 10)); m = 18.0 + 0.15 * t + rng.normal(scale=0.3, size=10); e = rng.uniform(0.02, 0.3, 10)
    e[rng.choice(10, size=2, replace=False)] = np.inf
    times[f], mags[f], errs[f] = t, m, e
limits = {"ztfg": 21.7, "ztfr": 21.4, "ztfi": 20.9, "sdssu": 23.9, "ps1::g": 25.0, "ps1::r": 24.7, "ps1::i": 24.0,
          "ps1::z": 23.3, "ps1::y": 22.1}
core = syn.random_model("Ka2017", gfilters, kind="gp", seed=2, Ntr=329)
priors = syn.ka2017_prior(); priors["timeshift"].maximum = 0.2
lik = build(core, "Ka2017", gfilters, (times, mags, errs, 0.0), priors, kind="gp", limit=limits)
rate("C4 Ka2017 sklearn_gp Ntr=329, 9 filters, detection limits", lik, priors, N_GP)
