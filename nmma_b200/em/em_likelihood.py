"""EM transient likelihood: reference API on top of the CUDA engine.

Mirror of ``nmma/em/em_likelihood.py`` (``EMTransientLikelihood`` :42-132,
``MultiFilterTransient`` :266-355) behind ``NMMALikelihood`` (``nmma/core/base.py:37-185``).
``log_likelihood(parameters: dict) -> float`` keeps the bilby contract (sentinel instead of
exceptions); ``log_likelihood_batch(points[N,P], columns)`` is the new batched entry point and
``pool`` a ``map``-compatible adapter for dynesty's ``pool=``/``queue_size=`` seam.

All arithmetic of a likelihood evaluation runs on the GPU: parameter conversion, redshift
lookup, surrogate, SVD reconstruction, both interpolation stages, systematics and the
(truncated) Gaussian / log-survival terms.  There is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .. import _lib as L
from .._lib import ParamSrc
from ..core.base import NMMALikelihood
from ..core.constants import SENTINEL
from ..core.priors import fixed_value, is_constraint, is_fixed_prior
from . import utils
from .model import resolve_param_sources


class MultiFilterTransient:
    """Multi-filter EM transient (``nmma/em/em_likelihood.py:266-355`` + base :136-263).

    Parameters follow the reference: ``filters`` (observed), ``light_curve_model``
    (:class:`~nmma_b200.em.model.SVDLightCurveModel`), ``light_curve_data`` =
    ``(times{f} - trigger, mags{f}, errs{f}, trigger_time)`` from ``setup_filtered_lc_data``,
    ``systematics_handler``, ``priors``, ``detection_limit``, ``verbose``.
    """

    def __init__(self, filters, light_curve_model, light_curve_data, systematics_handler, priors,
                 detection_limit=np.inf, verbose=False):
        self.observed_filters = list(filters)
        self.light_curve_model = light_curve_model
        self.model_filter_mapping, self.obs_average_mapping = utils.get_filter_name_mapping(
            self.observed_filters, extra_known=light_curve_model.filters)
        self.light_curve_model.check_vs_priors(priors)
        (self.light_curve_times, self.light_curves,
         self.light_curve_uncertainties, self.trigger_time) = light_curve_data
        systematics_handler.reset(self.light_curve_model.model_times, priors)
        self.systematics_handler = systematics_handler
        self.priors = priors
        self.verbose = verbose
        self.set_detection_limit(detection_limit)
        self._engine = None
        self._columns: Optional[List[str]] = None
        self._always_fail = False

    def set_detection_limit(self, detection_limit):
        self.detection_limit = utils.set_filter_associated_dict(detection_limit, self.observed_filters)
        self._engine = None

    def __repr__(self):
        return f"{self.__class__.__name__} (light_curve_model={self.light_curve_model})"

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_engine"] = None
        return state

    # ---- layout ---------------------------------------------------------------------------
    def default_columns(self) -> List[str]:
        """Sampled (non-fixed, non-constraint) prior keys in prior order."""
        return [k for k, p in self.priors.items() if not is_fixed_prior(p) and not is_constraint(p)]

    def _available(self, columns: Sequence[str]) -> Dict[str, ParamSrc]:
        avail = {k: ParamSrc.const(fixed_value(p)) for k, p in self.priors.items()
                 if is_fixed_prior(p) and not is_constraint(p)}
        for i, k in enumerate(columns):
            avail[k] = ParamSrc.column(i)
        return avail

    def _filter_index(self, name):
        model = self.light_curve_model
        if name not in model._eval_filters:
            raise KeyError(name)
        return model._eval_filters.index(name)

    def plan_layout(self, columns: Sequence[str]) -> dict:
        """Host-only: everything the engine needs for ``columns`` (no device access)."""
        model = self.light_curve_model
        avail = self._available(columns)
        plan = dict(P=len(columns))
        plan["xsrc"] = resolve_param_sources(model.model_parameters, avail)
        plan["dl"] = avail.get("luminosity_distance", ParamSrc.const(1e-5))
        plan["ts"] = avail.get("timeshift", ParamSrc.const(0.0))
        if "Ebv" in avail and not (avail["Ebv"].col < 0 and avail["Ebv"].value == 0.0):
            raise NotImplementedError("extinction (Ebv != 0) is not part of this build (DESIGN.md, 'next' rows)")
        plan["z_table"] = None
        if "redshift" in avail:
            plan["zsrc"], plan["zmode"] = avail["redshift"], L.Z_PARAM
        elif "luminosity_distance" in avail:
            if model._z_table is None:
                raise ValueError("luminosity_distance is sampled but not in the priors: the per-call "
                                 "z_at_value path of the reference is not available in batched mode")
            plan["z_table"] = model._z_table
            plan["zsrc"], plan["zmode"] = ParamSrc.const(0.0), L.Z_TABLE
        else:
            plan["zsrc"], plan["zmode"] = ParamSrc.const(0.0), L.Z_ZERO
        # a model filter the surrogate cannot evaluate is all-inf -> sanity_check fails for every point
        plan["always_fail"] = any(f not in model._eval_filters for f in model.filters)

        sysplan = self.systematics_handler.device_plan()
        obs_filters = [f for f in sysplan.keys()]          # band_log_likelihood iterates obs_error.items()
        helper_lists, times, mags, sigmas, limits = [], [], [], [], []
        modes, budgets, node_srcs, node_times = [], [], [], []
        for filt in obs_filters:
            if filt in self.model_filter_mapping:
                helper_lists.append([self._filter_index(self.model_filter_mapping[filt])])
            else:
                helper_lists.append([self._filter_index(self.model_filter_mapping[h])
                                     for h in self.obs_average_mapping[filt]])
            times.append(np.asarray(self.light_curve_times[filt], float))
            mags.append(np.asarray(self.light_curves[filt], float))
            sigmas.append(np.asarray(self.light_curve_uncertainties[filt], float))
            limits.append(float(self.detection_limit[filt]))
            entry = sysplan[filt]
            if entry[0] == "budget":
                modes.append(L.SYS_BUDGET); budgets.append(entry[1]); node_srcs.append([]); node_times.append([])
            elif entry[0] == "param":
                modes.append(L.SYS_PARAM); budgets.append(0.0)
                node_srcs.append([avail[entry[1]]]); node_times.append([0.0])
            else:
                modes.append(L.SYS_INTERP); budgets.append(0.0)
                node_srcs.append([avail[n] for n in entry[1]]); node_times.append(list(entry[2]))
        plan["obs"] = (helper_lists, times, mags, sigmas, limits)
        plan["sys"] = (modes, budgets, node_srcs, node_times)
        plan["obs_filters"] = obs_filters
        return plan

    def _build_engine(self, columns: Sequence[str]):
        model = self.light_curve_model
        plan = self.plan_layout(columns)
        eng = model.new_engine()
        eng.set_sample_grid(np.asarray(model.model_times, float))
        if plan["z_table"] is not None:
            eng.set_redshift_table(*plan["z_table"])
        eng.set_param_layout(plan["P"], plan["xsrc"], plan["dl"], plan["ts"], plan["zsrc"], plan["zmode"])
        eng.set_observations(*plan["obs"])
        eng.set_systematics(*plan["sys"])
        self._always_fail = plan["always_fail"]
        if hasattr(self.priors, "device_plan"):
            try:
                eng.set_priors(*self.priors.device_plan(columns))
            except (NotImplementedError, KeyError):
                pass        # a column without a device prior: prior_transform / sweep raise ERR_STATE when used
        self._engine = eng
        self._columns = list(columns)
        return eng

    def engine_for(self, columns: Optional[Sequence[str]] = None):
        columns = list(columns) if columns is not None else (self._columns or self.default_columns())
        if self._engine is None or self._columns != columns:
            self._build_engine(columns)
        return self._engine

    # ---- evaluation -------------------------------------------------------------------------
    def log_likelihood_batch(self, points, columns: Optional[Sequence[str]] = None, out=None):
        """log L for ``points[N, P]``.  NumPy in -> NumPy out (H2D/D2H inside the C call; page-locked
        arrays are copied without staging); CUDA tensor in -> CUDA tensor out (asynchronous on the
        current stream).  ``out`` optionally receives the result."""
        eng = self.engine_for(columns)
        is_tensor = hasattr(points, "is_cuda")
        if is_tensor:
            out = eng.logl_device(points, out=out)
            if self._always_fail:
                out.fill_(SENTINEL)
            return out
        out = eng.logl_host(points, out=out)
        if self._always_fail:
            out[:] = SENTINEL
        return out

    def prior_transform_batch(self, unit, columns: Optional[Sequence[str]] = None):
        """``PriorDict.rescale`` on the device: ``unit[N,P]`` in the unit cube -> CUDA ``points[N,P]``."""
        return self.engine_for(columns).prior_transform(unit)

    def sample_prior_batch(self, n, seed=0, first_index=0, columns: Optional[Sequence[str]] = None):
        """``n`` prior draws on the device (counter-based: draw ``first_index + i`` does not depend on sharding)."""
        return self.engine_for(columns).prior_sample(n, seed, first_index)

    def log_likelihood_sweep(self, n, seed=0, first_index=0, return_points=False,
                             columns: Optional[Sequence[str]] = None):
        """log L of ``n`` prior draws with no host traffic; sentinel semantics as ``log_likelihood_batch``."""
        eng = self.engine_for(columns)
        res = eng.logl_sweep(n, seed, first_index, return_points)
        if self._always_fail:
            (res[0] if return_points else res).fill_(SENTINEL)
        return res

    def log_likelihood(self, parameters):
        """One point, dict in / float out (``nmma/em/em_likelihood.py:186-204``)."""
        eng = self.engine_for(None)
        row = np.array([[parameters[k] for k in self._columns]], float)
        logl = float(eng.logl_host(row)[0])
        if self._always_fail:
            logl = SENTINEL
        if self.verbose:
            print(parameters, logl)
        return logl


class EMTransientLikelihood(NMMALikelihood):
    """Generic EM transient likelihood (``nmma/em/em_likelihood.py:42-132``), same signature."""

    def __init__(self, light_curve_model, light_curve_data, systematics_handler, priors, filters=None,
                 detection_limit=np.inf, verbose=False, **kwargs):
        if not filters:
            raise NotImplementedError("bolometric (filter-less) transients are outside the nmma_b200 hot path")
        sub_model = MultiFilterTransient(filters, light_curve_model, light_curve_data, systematics_handler,
                                         priors, detection_limit, verbose)
        super().__init__(sub_model, priors, **kwargs)

    def setup_submodel_conversion(self):
        self.conv_functions.append(self.sub_model.light_curve_model.parameter_conversion)

    def sanity_checks(self):
        return self.sub_model.light_curve_model.good_parameters

    def __repr__(self):
        return f"{self.__class__.__name__} based on {self.sub_model.__repr__()}"

    # ---- bilby contract -------------------------------------------------------------------------
    def log_likelihood(self, parameters=None):
        """``nmma/core/base.py:77-82``.  The conversion chain of the reference only adds derived keys
        (KNtheta, log10 twins); the device performs the same conversions from the sampled columns, so
        the dict is evaluated as given."""
        if parameters is None:
            parameters = self.parameters
        if self.constraints and not self.evaluate_constraints(self.parameter_conversion(dict(parameters))):
            return SENTINEL
        if not self.sanity_checks():
            return SENTINEL
        return self.sub_log_likelihood(parameters)

    # ---- batched entry points -------------------------------------------------------------------
    def log_likelihood_batch(self, points, columns: Optional[Sequence[str]] = None, out=None):
        """``float64[N]`` log L (or sentinel) for ``points[N, P]``; ``columns`` names the P columns
        (default: sampled prior keys in prior order)."""
        if self.constraints:
            raise NotImplementedError("Constraint priors are evaluated per point on the host; "
                                      "use log_likelihood(dict) or drop the constraint for batched sweeps")
        return self.sub_model.log_likelihood_batch(points, columns, out=out)

    def prior_transform_batch(self, unit, columns: Optional[Sequence[str]] = None):
        """Vectorised ``priors.rescale`` on the device (ultranest ``transform`` with ``vectorized=True``)."""
        return self.sub_model.prior_transform_batch(unit, columns)

    def sample_prior_batch(self, n, seed=0, first_index=0, columns: Optional[Sequence[str]] = None):
        return self.sub_model.sample_prior_batch(n, seed, first_index, columns)

    def log_likelihood_sweep(self, n, seed=0, first_index=0, return_points=False,
                             columns: Optional[Sequence[str]] = None):
        """Prior sweep (BASELINE.json configs[1] and [4]) with the draws made on the device."""
        if self.constraints:
            raise NotImplementedError("Constraint priors are evaluated per point on the host")
        return self.sub_model.log_likelihood_sweep(n, seed, first_index, return_points, columns)

    def vectorized(self, columns: Optional[Sequence[str]] = None):
        """``(transform, loglike)`` callables for vectorised samplers (ultranest ``ReactiveNestedSampler(
        param_names, loglike, transform, vectorized=True)``; dynesty ``prior_transform`` on stacked live points):
        NumPy ``[N,P]`` in, NumPy out, one GPU call each."""
        cols = list(columns) if columns is not None else self.columns

        def transform(unit):
            return self.prior_transform_batch(np.atleast_2d(np.asarray(unit, float)), cols).cpu().numpy()

        def loglike(theta):
            return self.log_likelihood_batch(np.atleast_2d(np.asarray(theta, float)), cols)

        return transform, loglike

    @property
    def columns(self):
        return self.sub_model.default_columns()

    @property
    def pool(self):
        return BatchPool(self)


OpticalLightCurve = None  # set in nmma_b200.em.likelihood (legacy positional signature)


class BatchPool:
    """``pool.map(fn, list_of_theta)`` adapter (dynesty / schwimmbad seam, ``nmma/core/mpi_setup.py:298-303,679``):
    stacks the points of one sampler iteration and makes a single GPU call; ``fn`` is ignored for
    likelihood calls issued by the sampler (it *is* this likelihood)."""

    def __init__(self, likelihood: EMTransientLikelihood, columns: Optional[Sequence[str]] = None):
        self.likelihood = likelihood
        self.columns = list(columns) if columns is not None else likelihood.columns
        self.size = 1 << 14

    def map(self, fn, iterable):
        thetas = list(iterable)
        if len(thetas) == 0:
            return []
        first = thetas[0]
        if isinstance(first, dict):
            pts = np.array([[t[k] for k in self.columns] for t in thetas], float)
        else:
            pts = np.asarray(thetas, float)
        if pts.ndim != 2 or pts.shape[1] != len(self.columns):
            return [fn(t) for t in thetas]          # not a likelihood call (e.g. prior transform)
        return list(self.likelihood.log_likelihood_batch(pts, self.columns))

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
