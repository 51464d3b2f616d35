"""Legacy module name ``nmma/em/likelihood.py`` (the north-star API).

``OpticalLightCurve(light_curve_model, filters, light_curve_data, trigger_time, error_budget=1,
tmin=0, tmax=14)`` is the signature that survives in the reference's stale
``nmma/em/__pycache__/likelihood.cpython-37.pyc``; in the current source the same job is done by
``EMTransientLikelihood`` (``nmma/em/em_likelihood.py:42-132``).  Both names are served by the
same GPU path.
"""
from __future__ import annotations

import numpy as np

from ..core.priors import PriorDict
from . import utils
from .em_likelihood import BatchPool, EMTransientLikelihood, MultiFilterTransient  # noqa: F401
from .systematics import FilterSystematicsHandler


def _standardise(light_curve_data, filters):
    """Old layout {filt: array[[t, mag, err], ...]} or new {filt: {'time','mag','mag_error'}}."""
    out = {}
    for filt in filters:
        if filt not in light_curve_data:
            continue
        d = light_curve_data[filt]
        if isinstance(d, dict):
            out[filt] = {k: np.asarray(d[k], float) for k in ("time", "mag", "mag_error")}
        else:
            a = np.asarray(d, float)
            out[filt] = {"time": a[:, 0], "mag": a[:, 1], "mag_error": a[:, 2]}
    return out


class OpticalLightCurve(EMTransientLikelihood):
    """Legacy optical kilonova likelihood; times in ``light_curve_data`` are absolute (MJD) and
    are cut to ``[tmin, tmax]`` days after ``trigger_time`` as the legacy class did."""

    def __init__(self, light_curve_model, filters, light_curve_data, trigger_time, detection_limit=None,
                 error_budget=1.0, tmin=0.0, tmax=14.0, verbose=False, priors=None, systematics_file=None):
        if isinstance(filters, str):
            filters = filters.split(",")
        data = _standardise(light_curve_data, filters)
        for filt in list(data):
            t = data[filt]["time"] - trigger_time
            keep = (t >= tmin) & (t <= tmax)
            if not keep.any():
                del data[filt]
            else:
                data[filt] = {k: v[keep] for k, v in data[filt].items()}
        filters = [f for f in filters if f in data]
        lc_data = utils.setup_filtered_lc_data(data, trigger_time)
        if priors is None:
            priors = PriorDict()
        handler = FilterSystematicsHandler(filters, systematics_file, error_budget, lc_data[0])
        if systematics_file is not None:
            handler.setup_systematics_priors(priors)
        if detection_limit is None:
            detection_limit = np.inf
        self._lazy_columns = len(priors) == 0
        super().__init__(light_curve_model, lc_data, handler, priors, filters=filters,
                         detection_limit=detection_limit, verbose=verbose)
        self.trigger_time = trigger_time
        self.tmin, self.tmax = tmin, tmax
        self.error_budget = error_budget

    def log_likelihood(self, parameters=None):
        if parameters is None:
            parameters = self.parameters
        if self._lazy_columns and self.sub_model._columns is None:
            # no priors given: every numeric key of the first call becomes a column
            cols = [k for k, v in parameters.items() if np.isscalar(v) and not isinstance(v, str)]
            self.sub_model.engine_for(cols)
        return super().log_likelihood(parameters)
