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
from .model import resolve_constraint_sources, resolve_param_sources


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
        self._engines: Dict[tuple, tuple] = {}       # column tuple -> (engine, always_fail); _engine is the last one used
        self._columns: Optional[List[str]] = None
        self._dict_base: Optional[List[str]] = None  # sampled columns of log_likelihood(dict) calls
        self._always_fail = False

    def set_detection_limit(self, detection_limit):
        self.detection_limit = utils.set_filter_associated_dict(detection_limit, self.observed_filters)
        self._engine = None

    def __repr__(self):
        return f"{self.__class__.__name__} (light_curve_model={self.light_curve_model})"

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_engine"] = None
        state["_engines"] = {}
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
        plan["ext"] = None
        if "Ebv" in avail and not (avail["Ebv"].col < 0 and avail["Ebv"].value == 0.0):
            law, nu0, coef = model.extinction_plan(model._eval_filters)
            plan["ext"] = (law, avail["Ebv"], nu0, coef)
        # Constraint priors (nmma/core/base.py:67-68) on a column, a fixed value or a derived key of the conversion chain
        cons = [(k, p) for k, p in self.priors.items() if is_constraint(p)]
        plan["constraints"] = None
        if cons:
            srcs = resolve_constraint_sources([k for k, _ in cons], model.model_parameters, avail)
            plan["constraints"] = (srcs, [float(p.minimum) for _, p in cons], [float(p.maximum) for _, p in cons])
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
        if plan["ext"] is not None:
            eng.set_extinction(*plan["ext"])
        if plan["constraints"] is not None:
            eng.set_constraints(*plan["constraints"])
        self._always_fail = plan["always_fail"]
        if hasattr(self.priors, "device_plan") and all(c in self.priors for c in columns):
            try:
                eng.set_priors(*self.priors.device_plan(columns))
            except NotImplementedError:
                pass        # a column whose prior has no device transform: prior_transform / sweep raise ERR_STATE when used
        self._engine = eng
        self._columns = list(columns)
        self._engines[tuple(columns)] = (eng, self._always_fail)
        return eng

    def engine_for(self, columns: Optional[Sequence[str]] = None):
        columns = list(columns) if columns is not None else (self._columns or self.default_columns())
        if self._engine is None:              # invalidated (new detection limit, unpickled, test hooks): drop every layout
            self._engines = {}
        if self._engine is None or self._columns != columns:
            hit = self._engines.get(tuple(columns))
            if hit is not None:
                self._engine, self._always_fail = hit
                self._columns = list(columns)
            else:
                self._build_engine(columns)
        return self._engine

    _ANGLE_KEYS = ("KNtheta", "inclination_EM", "theta_jn", "cos_theta_jn")

    def _dict_columns(self, parameters) -> List[str]:
        """Columns for ``log_likelihood(dict)``.  The reference evaluates the dict as given
        (``em_parameter_setup``, nmma/em/model.py:288-303), so besides the sampled keys a key becomes a column when
        its value is not the constant staged from the priors: a fixed prior given a different value, or a key
        without a prior that the path reads (``luminosity_distance``, ``timeshift``, ``Ebv``, ``redshift`` when no
        distance prior exists, a model parameter, the first angle key when none is sampled)."""
        cols = list(self._dict_base)
        pri = self.priors
        have_angle = any(a in cols or a in pri for a in self._ANGLE_KEYS)
        extra = []
        for k, v in parameters.items():
            if k in cols or isinstance(v, (str, bool)) or not np.isscalar(v):
                continue
            p = pri.get(k) if k in pri else None
            if p is not None:
                if is_constraint(p) or not is_fixed_prior(p) or fixed_value(p) == float(v):
                    continue
            elif k in self._ANGLE_KEYS:
                if have_angle:
                    continue          # a twin the conversion chain derived from the sampled angle
                have_angle = True
            elif k in ("luminosity_distance", "timeshift") or k in self.light_curve_model.model_parameters:
                pass
            elif k == "Ebv":
                if float(v) == 0.0:
                    continue
            elif k == "redshift":
                if "luminosity_distance" in pri or "luminosity_distance" in cols:
                    continue          # check_vs_priors installed the dL -> z table: the dict's redshift is not read
            else:
                continue
            extra.append(k)
        return cols + extra

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
        if self._dict_base is None:
            self._dict_base = list(self._columns or self.default_columns())
        cols = self._dict_columns(parameters)
        eng = self.engine_for(cols)
        row = np.array([[parameters[k] for k in cols]], float)
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
    def posterior_conversion(self, posterior_samples):
        """Derived posterior columns (``nmma/em/em_likelihood.py:124-132``): total ejecta mass, GRB wing angle twins.
        Works on anything indexable by key with array values (a pandas DataFrame, a dict of arrays)."""
        ps = posterior_samples
        if "log10_mej_dyn" in ps and "log10_mej_wind" in ps:
            ps["log10_mej"] = np.log10(10 ** (ps["log10_mej_wind"]) + 10 ** (ps["log10_mej_dyn"]))
        if "thetaWing" in ps and "thetaCore" in ps:
            ps["alphaWing"] = ps["thetaWing"] / ps["thetaCore"]
        elif "alphaWing" in ps and "thetaCore" in ps:
            ps["thetaWing"] = ps["alphaWing"] * ps["thetaCore"]
        return ps

    def final_diagnostics(self, bestfit_params, args=None, result=None):
        """The reference plots the best-fit light curve here (matplotlib, out of scope); this returns what the plot shows:
        ``(observable_times, {filter: apparent magnitudes})`` of the best fit, evaluated on the GPU."""
        params = self.parameter_conversion(dict(bestfit_params))
        return self.sub_model.light_curve_model.gen_detector_lc(params)

    def log_likelihood(self, parameters=None):
        """``nmma/core/base.py:77-82``.  The conversion chain of the reference only adds derived keys
        (KNtheta, log10 twins); the device performs the same conversions from the sampled columns, so
        the dict is evaluated as given."""
        if parameters is None:
            parameters = self.parameters
        if not self.sanity_checks():
            return SENTINEL
        return self.sub_log_likelihood(parameters)

    # ---- batched entry points -------------------------------------------------------------------
    def log_likelihood_batch(self, points, columns: Optional[Sequence[str]] = None, out=None):
        """``float64[N]`` log L (or sentinel) for ``points[N, P]``; ``columns`` names the P columns
        (default: sampled prior keys in prior order)."""
        return self.sub_model.log_likelihood_batch(points, columns, out=out)   # Constraint priors: on the device

    def prior_transform_batch(self, unit, columns: Optional[Sequence[str]] = None):
        """Vectorised ``priors.rescale`` on the device (ultranest ``transform`` with ``vectorized=True``)."""
        return self.sub_model.prior_transform_batch(unit, columns)

    def sample_prior_batch(self, n, seed=0, first_index=0, columns: Optional[Sequence[str]] = None):
        return self.sub_model.sample_prior_batch(n, seed, first_index, columns)

    def log_likelihood_sweep(self, n, seed=0, first_index=0, return_points=False,
                             columns: Optional[Sequence[str]] = None):
        """Prior sweep (BASELINE.json configs[1] and [4]) with the draws made on the device."""
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
        if getattr(self, "_pool", None) is None:
            self._pool = BatchPool(self)
        return self._pool


OpticalLightCurve = None  # set in nmma_b200.em.likelihood (legacy positional signature)


class BatchPool:
    """``pool.map(fn, iterable)`` adapter for the dynesty / schwimmbad seam
    (``nmma/core/mpi_setup.py:298-303,679``; dynesty calls ``pool.map(loglikelihood, points)`` and
    ``pool.map(prior_transform, unit_points)`` with ``queue_size`` points per call).

    Hand ``pool.loglike`` (and optionally ``pool.prior_transform``) to the sampler and pass the pool as
    ``pool=``: a ``map`` whose ``fn`` is one of these (directly, as a bound method of the likelihood, or inside the
    sampler's wrapper object) stacks the points of the call and makes a single GPU call.  Any other ``fn``
    (``evolve_point`` with argument objects, user callbacks) is mapped serially, untouched.
    """

    _UNWRAP = ("func", "loglikelihood", "prior_transform", "__wrapped__", "__func__")

    def __init__(self, likelihood: "EMTransientLikelihood", columns: Optional[Sequence[str]] = None):
        self.likelihood = likelihood
        self.columns = list(columns) if columns is not None else likelihood.columns
        self.size = 1 << 14

    # ---- the callables to give to the sampler -------------------------------------------------------------
    def loglike(self, theta):
        """One point (array in ``columns`` order, or dict) -> float log L."""
        if isinstance(theta, dict):
            return float(self.likelihood.log_likelihood(theta))
        row = np.asarray(theta, float).reshape(1, -1)
        return float(self.likelihood.log_likelihood_batch(row, self.columns)[0])

    def prior_transform(self, unit):
        """One unit-cube point -> physical point (``priors.rescale``), evaluated on the device."""
        row = np.asarray(unit, float).reshape(1, -1)
        return self.likelihood.prior_transform_batch(row, self.columns).cpu().numpy()[0]

    # ---- dispatch -----------------------------------------------------------------------------------------
    def _role(self, fn):
        """'loglike' / 'prior' when ``fn`` is (a wrapper around) one of this pool's or likelihood's callables."""
        lik = self.likelihood
        seen = 0
        while fn is not None and seen < 6:
            owner = getattr(fn, "__self__", None)
            name = getattr(fn, "__name__", "")
            if (owner is self or owner is lik or owner is getattr(lik, "sub_model", None)
                    or (isinstance(owner, BatchPool) and owner.likelihood is lik)):
                if name in ("loglike", "log_likelihood", "log_likelihood_ratio"):
                    return "loglike"
                if name == "prior_transform":
                    return "prior"
            nxt = None
            for attr in self._UNWRAP:
                cand = getattr(fn, attr, None)
                if callable(cand) and cand is not fn:
                    nxt = cand
                    break
            fn, seen = nxt, seen + 1
        return None

    def map(self, fn, iterable):
        items = list(iterable)
        if not items:
            return []
        role = self._role(fn)
        if role is None:
            return [fn(t) for t in items]
        try:
            if role == "loglike" and isinstance(items[0], dict):
                pts = np.array([[t[k] for k in self.columns] for t in items], float)
            else:
                pts = np.asarray(items, float)
        except (TypeError, ValueError, KeyError):
            return [fn(t) for t in items]
        if pts.ndim != 2 or pts.shape[1] != len(self.columns):
            return [fn(t) for t in items]
        if role == "prior":
            return list(self.likelihood.prior_transform_batch(pts, self.columns).cpu().numpy())
        return [float(v) for v in self.likelihood.log_likelihood_batch(pts, self.columns)]

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
