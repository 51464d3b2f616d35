"""TEST INFRASTRUCTURE ONLY -- never imported by the product (nmma_b200/).

Builds the oracle-side objects (``oracle.nmma_oracle``) from the same in-memory
configuration the GPU likelihood is built from, so that tests, ``smoke()`` and
``bench.py``'s CPU legs evaluate identical inputs through both implementations.
"""
from __future__ import annotations

import copy

import numpy as np

from . import nmma_oracle as O


def oracle_core(core):
    """Reference in-memory layout with Keras stand-ins for the (W1, b1, W2, b2) tuples."""
    out = {}
    for filt, entry in core.items():
        e = dict(entry)
        if "model" in e and isinstance(e["model"], (tuple, list)):
            e["model"] = O.KerasStandIn(*e["model"])
        out[filt] = e
    return out


def is_fixed(prior):
    return isinstance(prior, (int, float)) or hasattr(prior, "peak")


def build_oracle_likelihood(core, model_parameters, model_filters, sample_times, obs_filters,
                            light_curve_data, priors, sys_plan=None, error_budget=1.0,
                            detection_limit=np.inf, mag_ncoeff=None, z_table=None, filts_lambdas=None,
                            extinction_law=None, extinction_coef=None):
    """Returns (oracle likelihood, fixed-parameter dict).

    ``sys_plan``: output of ``FilterSystematicsHandler.device_plan()`` (host-side map construction is
    not part of the per-point path) or None for a constant ``error_budget``.
    """
    ocore = oracle_core(core)
    dfilts, lambdas = filts_lambdas if filts_lambdas is not None else (None, None)
    model = O.OracleSVDLightCurveModel(model_parameters, ocore, filters=model_filters,
                                       sample_times=sample_times, mag_ncoeff=mag_ncoeff, default_filts=dfilts,
                                       lambdas=lambdas, extinction_coef=extinction_coef)
    if extinction_law is not None:
        model.extinction_law = extinction_law
    is_con = lambda p: p.__class__.__name__ == "Constraint"  # noqa: E731
    constraints = {k: (p.minimum, p.maximum) for k, p in priors.items() if is_con(p)}
    fixed = {k: float(getattr(p, "peak", p)) for k, p in priors.items() if is_fixed(p) and not is_con(p)}
    if "redshift" not in priors and "luminosity_distance" in priors:
        dl = priors["luminosity_distance"]
        if z_table is not None:
            model.check_vs_priors(table=z_table)
        else:
            model.check_vs_priors(dl.minimum, dl.maximum)
    times = light_curve_data[0]
    direct, interp, budget = {}, {}, error_budget
    if sys_plan is not None:
        budget = {}
        for f, entry in sys_plan.items():
            if entry[0] == "budget":
                budget[f] = entry[1]
            elif entry[0] == "param":
                direct[f] = entry[1]
            else:
                interp[f] = (entry[1], entry[2])
        if not budget:
            budget = 1.0
        else:
            for f in obs_filters:
                budget.setdefault(f, 1.0)
    sys_filters = list(sys_plan.keys()) if sys_plan is not None else list(obs_filters)
    sysh = O.OracleFilterSystematics(sys_filters, times, budget=budget, direct=direct, interp=interp)
    lik = O.OracleMultiFilterTransient(obs_filters, model, light_curve_data, sysh,
                                       detection_limit=detection_limit, constraints=constraints)
    return lik, fixed


class UnpackedGP:
    """Oracle-side stand-in for a fitted sklearn GaussianProcessRegressor rebuilt from unpacked arrays:
    ``predict`` follows sklearn's ``kernel_(X, X_train_) @ alpha_`` with RationalQuadratic.__call__
    (cdist(X/l, Y/l, 'sqeuclidean'); base = 1 + d/(2 alpha); K = C^2 * base**-alpha)."""

    def __init__(self, X, alpha, c2, ra, rl, ym, ys):
        self.X, self.alpha, self.c2, self.ra, self.rl, self.ym, self.ys = X, alpha, c2, ra, rl, ym, ys

    def predict(self, x, return_std=False):
        from scipy.spatial.distance import cdist
        d = cdist(np.atleast_2d(x) / self.rl, self.X / self.rl, metric="sqeuclidean")
        K = self.c2 * (1 + d / (2 * self.ra)) ** (-self.ra)
        y = self.ys * (K @ self.alpha) + self.ym
        return (y, np.zeros_like(y)) if return_std else y


def oracle_ready_core(core):
    """Replace unpacked GP dicts by objects with a ``predict`` method for the oracle."""
    out = {}
    for f, e in core.items():
        e = dict(e)
        if isinstance(e.get("gps"), dict):
            g = e["gps"]
            e["gps"] = [UnpackedGP(g["X"], g["alpha"][i], g["c2"][i], g["rq_alpha"][i], g["rq_len"][i],
                                   g["ymean"][i], g["ystd"][i]) for i in range(g["alpha"].shape[0])]
        out[f] = e
    return out


def oracle_logl(lik, fixed, points, columns):
    out = np.empty(len(points))
    for i, row in enumerate(np.asarray(points, float)):
        params = dict(fixed)
        params.update({k: float(v) for k, v in zip(columns, row)})
        out[i] = lik.log_likelihood(params)
    return out
