"""Workload builders for the benchmark / parity configurations.

The Zenodo/GitLab surrogate weights for Bu2019lm, Bu2023Ye and Ka2017 are not available
offline; BASELINE.json's north star allows random-init weights of the same architecture and
basis shape.  The photometry is the real AT2017gfo table (``nmma_b200/data/at2017gfo.json``).
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import numpy as np

from .core.priors import PriorDict, Sine, Uniform
from .em import utils
from .mlmodel import random_surrogate

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
AT2017GFO_TRIGGER_MJD = 57982.5285236896   # doc/training.md:79
AT2017GFO_FILTERS = ["ps1::g", "ps1::r", "ps1::i", "ps1::z", "ps1::y", "2massj", "2massh", "2massks", "sdssu"]

# training-grid bounds of the surrogates (priors/Bu2019lm.prior, Bu2023Ye.prior, Ka2017.prior)
GRID_BOUNDS = {
    "Bu2019lm": ([-3.0, -2.0, 0.0, 0.0], [-1.7, -0.89, 90.0, 90.0]),
    "Bu2023Ye": ([-3.0, 0.12, 0.15, -2.0, 0.03, 0.20, 0.0], [-1.7, 0.25, 0.30, -0.89, 0.15, 0.40, 90.0]),
    "Ka2017": ([-3.0, -1.52, -9.0], [-1.0, -0.53, -1.0]),
    "Bu2019nsbh": ([-2.0, -2.0, 0.0], [-1.04575749, -1.04575749, 90.0]),
}


def load_at2017gfo(data_tmax: float = 14.0, data_tmin: float = 0.0, filters=None):
    """(light_curve_data tuple as setup_filtered_lc_data returns it, filters)."""
    with open(os.path.join(DATA_DIR, "at2017gfo.json")) as fh:
        blob = json.load(fh)
    data = {f: {k: np.array([np.inf if x == "inf" else x for x in v], float) for k, v in d.items()}
            for f, d in blob["data"].items()}
    if filters is not None:
        data = {f: data[f] for f in filters if f in data}
    args = SimpleNamespace(data_tmin=data_tmin, data_tmax=data_tmax)
    data = utils.cut_data_to_time_range(data, args, AT2017GFO_TRIGGER_MJD)
    filters = [f for f in (filters or AT2017GFO_FILTERS) if f in data]
    return utils.setup_filtered_lc_data({f: data[f] for f in filters}, AT2017GFO_TRIGGER_MJD), filters


def random_model(name: str, filters, kind: str = "mlp", seed: int = 0, **kw):
    """Random-init surrogate of the named model's architecture in the reference's in-memory layout."""
    from .em.model import model_parameters_dict
    d = len(model_parameters_dict[name])
    mins, maxs = GRID_BOUNDS[name]
    return random_surrogate(filters, d=d, kind=kind, seed=seed, param_mins=mins, param_maxs=maxs, **kw)


def bu2019lm_prior() -> PriorDict:
    """``priors/Bu2019lm.prior`` (keys on the left-hand side win)."""
    p = PriorDict()
    p["luminosity_distance"] = Uniform(1, 200.0, name="luminosity_distance")
    p["KNphi"] = Uniform(15.0, 75.0, name="KNphi")
    p["inclination_EM"] = Sine(0.0, np.pi / 2.0, name="inclination_EM")
    p["timeshift"] = Uniform(-2.0, 0.1, name="timeshift")
    p["log10_mej_dyn"] = Uniform(-3.0, -1.0, name="log10_mej_dyn")
    p["log10_mej_wind"] = Uniform(-3.0, -0.5, name="log10_mej_wind")
    return p


def bu2023ye_prior() -> PriorDict:
    """``priors/Bu2023Ye.prior`` with the luminosity_distance minimum raised to 1 Mpc."""
    p = PriorDict()
    p["log10_mej_dyn"] = Uniform(-3.0, -1.7, name="log10_mej_dyn")
    p["vej_dyn"] = Uniform(0.12, 0.25, name="vej_dyn")
    p["Yedyn"] = Uniform(0.15, 0.30, name="Yedyn")
    p["log10_mej_wind"] = Uniform(-2.0, -0.89, name="log10_mej_wind")
    p["vej_wind"] = Uniform(0.03, 0.15, name="vej_wind")
    p["Yewind"] = Uniform(0.20, 0.40, name="Yewind")
    p["inclination_EM"] = Sine(0.0, np.pi / 2.0, name="inclination_EM")
    p["luminosity_distance"] = Uniform(1.0, 200.0, name="luminosity_distance")
    p["timeshift"] = Uniform(-2.0, 1.0, name="timeshift")
    return p


def ka2017_prior() -> PriorDict:
    """``priors/Ka2017.prior`` with the luminosity_distance minimum raised to 1 Mpc."""
    p = PriorDict()
    p["luminosity_distance"] = Uniform(1.0, 200.0, name="luminosity_distance")
    p["timeshift"] = Uniform(-2.0, 1.0, name="timeshift")
    p["log10_mej"] = Uniform(-3.0, -1.0, name="log10_mej")
    p["log10_vej"] = Uniform(-1.52, -0.53, name="log10_vej")
    p["log10_Xlan"] = Uniform(-9, -1, name="log10_Xlan")
    return p
