"""Systematic error budget: YAML -> priors and per-filter evaluation plan.

Mirror of ``nmma/em/systematics.py``.  Map construction, prior naming and the legacy
``config: withTime/withoutTime`` layout follow the reference (``:14-336``, ``:343-512``);
the per-point evaluation (``__call__``, ``:54,279-296``) runs on the GPU -- the handler only
describes it through :meth:`FilterSystematicsHandler.device_plan`.  ``__call__`` is kept for
API parity with the reference (diagnostics / plotting) and is never used by the likelihood.
"""
from __future__ import annotations

import inspect
import os
import warnings
from ast import literal_eval
from pathlib import Path

import numpy as np
import yaml

from ..core import priors as bprior
from .utils import autocomplete_data, set_filter_associated_dict


def load_yaml(file_path):
    """``nmma/core/utils.py:45-46``."""
    return yaml.safe_load(os.path.expandvars(Path(file_path).read_text()))


class ValidationError(ValueError):
    def __init__(self, key, message):
        super().__init__(f"Validation error for '{key}': {message}")


class SystematicsHandler:
    """``nmma/em/systematics.py:14-192`` (single light curve / bolometric case)."""
    allowed_keys = ["time_range", "time_nodes", "prior", "params", "each", "filters"]

    def __init__(self, systematics_file=None, error_budget=None, light_curve_times=np.linspace(0.1, 14, 10),
                 base_prior_name="em_syserr"):
        self.base_prior_name = base_prior_name
        self.default_t_grid_type = "linear"
        self.light_curve_times = light_curve_times
        self.adjust_error_budget(error_budget)
        self.compute_em_err = self.from_budget
        if isinstance(systematics_file, str):
            self.systematics_dict = load_yaml(systematics_file)
        elif isinstance(systematics_file, dict):
            self.systematics_dict = systematics_file
        else:
            self.systematics_dict = {}

    # -- evaluation (host mirror) -------------------------------------------------------------
    def adjust_error_budget(self, error_budget):
        if error_budget is None:
            error_budget = 0.0001
        if isinstance(error_budget, str):
            error_budget = float(error_budget)
        self.error_budget = np.full_like(self.light_curve_times, error_budget)

    def from_budget(self, _):
        return self.error_budget

    def __call__(self, parameters):
        return self.compute_em_err(parameters)

    def from_param(self, parameters):
        return np.full_like(self.light_curve_times, parameters[self.base_prior_name])

    def from_parameters(self, parameters):
        vals = [parameters[p] for p in self.err_params]
        return autocomplete_data(self.light_curve_times, self.time_nodes, vals, extrapolate="constant")

    # -- YAML -> priors -----------------------------------------------------------------------
    def setup_systematics_priors(self, prior_dict):
        """``:57-84``: add the ``em_syserr*`` priors the YAML asks for to ``prior_dict``."""
        for key, info in self.systematics_dict.items():
            if key == "config":
                return self.legacy_prior_setup(prior_dict)
            if key in self.allowed_keys:  # one global systematic uncertainty
                prior_dict.update(self.setup_filt_prior("", self.systematics_dict))
                return prior_dict
            new_priors = self.setup_filt_prior(key, info)
            targets = info.get("each", [key]) if isinstance(info, dict) else [key]
            for filt in targets:
                for k, v in new_priors.items():
                    new_key = k.replace(key, filt)
                    pv = self._clone(v)
                    pv.name = new_key
                    prior_dict[new_key] = pv
        return prior_dict

    @staticmethod
    def _clone(prior):
        import copy
        return copy.copy(prior)

    def setup_filt_prior(self, key, info):
        name = self.prior_name(key)
        if isinstance(info, float):
            return {name: bprior.DeltaFunction(name=name, peak=info)}
        num = info.get("time_nodes", info.get("time_range", "1").split()[-1])
        if int(num) >= 2:
            return {n: self.get_prior(info, n) for n in (f"{name}_{i}" for i in range(int(num)))}
        return {name: self.get_prior(info, name)}

    def get_prior(self, info, n):
        prior_str = info["prior"]
        cls = prior_str.split("(")[0]
        args = "(".join(prior_str.split("(")[1:])[:-1]
        if args:
            p = bprior.prior_from_string(f"{cls}({args})", n)
            p.name = n
            return p
        klass = getattr(bprior, cls)
        return klass(**info.get("params", {}), **info.get("kwargs", {}), name=n)

    def get_name_and_times(self, key, info):
        return self.prior_name(key), self.get_time_range(info)

    def prior_name(self, key):
        return f"{self.base_prior_name}_{key}" if key else self.base_prior_name

    def get_time_range(self, info):
        """``:123-149``: ``time_nodes: n`` over the model range or ``time_range: "[lin|log] t0 t1 n"``."""
        if not isinstance(info, dict):
            return None
        num = info.get("time_nodes", None)
        t_range = info.get("time_range", "").split()
        if num is None and t_range:
            num = t_range.pop(-1)
        if num is None:
            return None
        grid_type = self.default_t_grid_type
        if len(t_range) == 3:
            grid_type, t_start, t_end = t_range
        elif len(t_range) == 2:
            t_start, t_end = t_range
            try:
                float(t_start)
            except ValueError:
                grid_type, t_end = t_range
                t_start = self.time_range[0]
        elif len(t_range) == 0:
            t_start, t_end = self.time_range
        else:
            raise ValueError("time range specfication invalid")
        if "lin" in grid_type:
            return np.linspace(float(t_start), float(t_end), int(num))
        if ("log" in grid_type) or ("geo" in grid_type):
            return np.geomspace(float(t_start), float(t_end), int(num))

    def setup_systematics_sampling(self, priors):
        name, time_range = self.get_name_and_times("", self.systematics_dict)
        if time_range is None:
            self.base_prior_name = name
            assert name in priors, "Required systematics prior missing"
            self.compute_em_err = self.from_param
        else:
            self.err_params = [f"{name}_{i}" for i, _ in enumerate(time_range)]
            assert all(p in priors for p in self.err_params), "Required systematics prior missing"
            self.time_nodes = time_range
            self.compute_em_err = self.from_parameters

    def legacy_prior_setup(self, prior_dict):
        """``:163-180``: legacy ``config:`` YAML -> prior strings -> priors."""
        add = {}
        for line in get_prior_strings(self.systematics_dict):
            key, _, val = line.partition("=")
            add[key.replace(" ", "")] = val.strip()
        prior_dict.update(bprior.PriorDict(add))
        return prior_dict

    def reset(self, model_times, priors):
        """``:186-192``."""
        self.time_range = (model_times[0], model_times[-1])
        if self.systematics_dict:
            self.setup_systematics_sampling(priors)
        elif self.base_prior_name in priors:
            self.compute_em_err = self.from_param


class FilterSystematicsHandler(SystematicsHandler):
    """``nmma/em/systematics.py:194-336``: per-filter systematics."""

    def __init__(self, filters, systematics_file=None, error_budget=None, light_curve_times=np.linspace(0.1, 14, 10),
                 base_prior_name="em_syserr"):
        self.filters = filters
        if not isinstance(light_curve_times, dict):
            light_curve_times = {filt: light_curve_times for filt in filters}
        self.direct_sys_map = {}
        self.interpolate_map = {}
        super().__init__(systematics_file, error_budget, light_curve_times, base_prior_name)

    def adjust_error_budget(self, error_budget):
        if error_budget is None:
            error_budget = 1.0
        elif isinstance(error_budget, str):
            error_budget = literal_eval(error_budget)
        budget = set_filter_associated_dict(error_budget, self.filters, 1.0)
        self.budget_values = budget
        self.error_budget = {f: np.full_like(self.light_curve_times[f], budget[f]) for f in self.filters}

    def setup_systematics_sampling(self, priors):
        """``:212-263``: which prior(s) drive which filter."""
        self.direct_sys_map = {}
        self.interpolate_map = {}
        self.missing_filters = set(self.filters)
        cleared = False
        for key, info in self.systematics_dict.items():
            if key == "config":
                self.legacy_systematics_setup(self.systematics_dict)
                break
            elif key in self.allowed_keys:
                name, trange = self.get_name_and_times("", self.systematics_dict)
                for filt in self.filters:
                    self.check_names_and_times(filt, trange, name, priors)
                break
            elif key in self.filters:
                name, trange = self.get_name_and_times(key, info)
                self.check_names_and_times(key, trange, name, priors)
            elif isinstance(info, dict) and "filters" in info:
                name, trange = self.get_name_and_times(key, info)
                for filt in info["filters"]:
                    self.check_names_and_times(filt, trange, name, priors)
            elif isinstance(info, dict) and "each" in info:
                name, trange = self.get_name_and_times(key, info)
                for filt in info["each"]:
                    self.check_names_and_times(filt, trange, name.replace(key, filt), priors)
            else:
                cleared = True
                name, trange = self.get_name_and_times(key, info)
                for filt in self.missing_filters:
                    self.check_names_and_times(filt, trange, name, priors, clean=False)
        assert cleared or len(self.missing_filters) == 0, \
            f"Some filters are missing systematic uncertainty definitions: {self.missing_filters}"
        if not self.interpolate_map:
            if len(set(self.direct_sys_map.values())) == 1:
                self.compute_em_err = self.from_param
            else:
                self.compute_em_err = self.from_single_params
        elif not self.direct_sys_map:
            self.compute_em_err = self.from_interpolated_params
        else:
            self.compute_em_err = self.from_parameters

    def check_names_and_times(self, filt, time_range, prior_name, priors, clean=True):
        if clean:
            self.direct_sys_map.pop(filt, None)
            self.interpolate_map.pop(filt, None)
            self.missing_filters.remove(filt)
        if time_range is None:
            assert prior_name in priors, "Required systematics prior missing"
            self.direct_sys_map[filt] = prior_name
        else:
            names = [f"{prior_name}_{i}" for i, _ in enumerate(time_range)]
            for p in names:
                assert p in priors, f"Required systematics prior missing: {p}"
            self.interpolate_map[filt] = (names, time_range)

    # host mirrors of :279-296
    def from_param(self, parameters):
        em_err = parameters[self.base_prior_name]
        return {f: np.full_like(self.light_curve_times[f], em_err) for f in self.filters}

    def from_single_params(self, parameters):
        return {f: np.full_like(self.light_curve_times[f], parameters[p]) for f, p in self.direct_sys_map.items()}

    def from_interpolated_params(self, parameters):
        return {f: autocomplete_data(self.light_curve_times[f], nodes, [parameters[p] for p in names],
                                     extrapolate="constant")
                for f, (names, nodes) in self.interpolate_map.items()}

    def from_parameters(self, parameters):
        out = self.from_single_params(parameters)
        out.update(self.from_interpolated_params(parameters))
        return out

    def legacy_systematics_setup(self, systematics_dict):
        """``:298-336`` including the reference quirk that a ``null`` entry sends *every*
        filter to the group ``all`` and stops."""
        validate_only_one_true(systematics_dict)
        tdep = systematics_dict["config"]["withTime"]
        if not tdep["value"]:
            self.direct_sys_map = {f: self.base_prior_name for f in self.filters}
            self.missing_filters = set()
            return
        yaml_filters = list(tdep["filters"])
        validate_filters(yaml_filters)
        groups = {}
        for fg in yaml_filters:
            if fg is None:
                groups = {f: "all" for f in self.filters}
                self.missing_filters = set()
                break
            elif isinstance(fg, list):
                for f in fg:
                    self.missing_filters.remove(f)
                    groups[f] = "___".join(fg)
            else:
                groups[fg] = fg
                self.missing_filters.remove(fg)
        nodes = np.round(np.linspace(*self.time_range, tdep["time_nodes"]), decimals=2)
        self.interpolate_map = {f: ([f"{self.base_prior_name}_{name}_{i}" for i, _ in enumerate(nodes)], nodes)
                                for f, name in groups.items()}

    # -- what the GPU evaluates ----------------------------------------------------------------
    def device_plan(self):
        """Per filter: ('budget', value) | ('param', name) | ('interp', names, nodes).

        Selection follows ``compute_em_err`` of the reference: ``from_budget`` (no YAML, no sampled
        ``em_syserr``), ``from_param`` (one value for all filters) or the per-filter maps.  Filters the
        chosen branch does not return are left out, exactly as ``band_log_likelihood`` iterates
        ``obs_error.items()`` (``nmma/em/em_likelihood.py:340``)."""
        fn = getattr(self.compute_em_err, "__func__", None)
        cls = type(self)
        plan = {}
        if fn is cls.from_budget or fn is SystematicsHandler.from_budget:
            for f in self.filters:
                plan[f] = ("budget", float(self.budget_values[f]))
        elif fn is cls.from_param:
            for f in self.filters:
                plan[f] = ("param", self.base_prior_name)
        else:
            if fn in (cls.from_single_params, cls.from_parameters):
                for f, p in self.direct_sys_map.items():
                    plan[f] = ("param", p)
            if fn in (cls.from_interpolated_params, cls.from_parameters):
                for f, (names, nodes) in self.interpolate_map.items():
                    plan[f] = ("interp", list(names), np.asarray(nodes, float))
        return plan


# ------------------------------------------------------------------------------------------
# Legacy YAML validation + prior-string generation (``nmma/em/systematics.py:343-512``)
# ------------------------------------------------------------------------------------------
ALLOWED_FILTERS = [
    "2massh", "2massj", "2massks", "atlasc", "atlaso", "bessellb", "besselli", "bessellr", "bessellux",
    "bessellv", "ps1::g", "ps1::i", "ps1::r", "ps1::y", "ps1::z", "sdssu", "uvot::b", "uvot::u",
    "uvot::uvm2", "uvot::uvw1", "uvot::uvw2", "uvot::v", "uvot::white", "ztfg", "ztfi", "ztfr",
]

ALLOWED_DISTRIBUTIONS = {n: c for n, c in inspect.getmembers(bprior, inspect.isclass)
                         if issubclass(c, bprior.Prior) and c is not bprior.Prior}


def get_positional_args(cls):
    sig = inspect.signature(cls.__init__)
    return [p.name for p in sig.parameters.values() if p.name != "self" and p.default == inspect.Parameter.empty]


DISTRIBUTION_PARAMETERS = {k: get_positional_args(v) for k, v in ALLOWED_DISTRIBUTIONS.items()}
# priors/systematics.yaml ships `min`/`max`; the validator wants `minimum`/`maximum` -- accept both
_KEY_ALIASES = {"min": "minimum", "max": "maximum"}


def validate_only_one_true(yaml_dict):
    for key, values in yaml_dict["config"].items():
        if "value" not in values or not isinstance(values["value"], bool):
            raise ValidationError(key, "'value' key must be present and be a boolean")
    n_true = sum(v["value"] for v in yaml_dict["config"].values())
    if n_true > 1:
        raise ValidationError("config", "Only one configuration key can be set to True at a time")
    if n_true == 0:
        raise ValidationError("config", "At least one configuration key must be set to True")


def validate_filters(filter_groups):
    allowed = ", ".join(str(f) for f in ALLOWED_FILTERS)
    used = set()
    for group in filter_groups:
        if isinstance(group, list):
            in_group = set()
            for filt in group:
                if filt not in ALLOWED_FILTERS:
                    raise ValidationError("filters", f"Invalid filter value '{filt}'. Allowed values are {allowed}")
                if filt in in_group:
                    raise ValidationError("filters", f"Duplicate filter value '{filt}' within the same group.")
                if filt in used:
                    raise ValidationError("filters", f"Duplicate filter value '{filt}'. A filter can only be used in one group.")
                used.add(filt)
                in_group.add(filt)
        elif group is not None and group not in ALLOWED_FILTERS:
            raise ValidationError("filters", f"Invalid filter value '{group}'. Allowed values are {allowed}")
        elif group in used:
            raise ValidationError("filters", f"Duplicate filter value '{group}'. A filter can only be used in one group.")
        else:
            used.add(group)


def _normalised(distribution):
    return {_KEY_ALIASES.get(k, k): v for k, v in distribution.items()}


def validate_distribution(distribution):
    distribution = _normalised(distribution)
    dist_type = distribution.get("type")
    if dist_type not in ALLOWED_DISTRIBUTIONS:
        raise ValidationError("distribution type",
                              f"Invalid distribution '{dist_type}'. Allowed values are {', '.join(str(f) for f in ALLOWED_DISTRIBUTIONS)}")
    missing = set(DISTRIBUTION_PARAMETERS[dist_type]) - set(distribution.keys())
    if missing:
        raise ValidationError("distribution", f"Missing required parameters for {dist_type} distribution: {', '.join(missing)}")


def create_prior_string(name, distribution):
    distribution = _normalised(distribution)
    dist_type = distribution["type"]
    klass = ALLOWED_DISTRIBUTIONS[dist_type]
    required = DISTRIBUTION_PARAMETERS[dist_type]
    params = {k: v for k, v in distribution.items() if k not in ["type", "value", "time_nodes", "filters"]}
    extra = set(params.keys()) - set(required)
    if extra:
        warnings.warn(f"Distribution parameters {extra} are not used by {dist_type} distribution and will be ignored")
    params = {k: params[k] for k in required if k in params}
    return f"{name} = {repr(klass(**params, name=name))}"


def handle_withTime(values):
    validate_distribution(values)
    groups = values.get("filters", [])
    validate_filters(groups)
    out = []
    for group in groups:
        if isinstance(group, list):
            gname = "___".join(group)
        else:
            gname = group if group is not None else "all"
        for n in range(values["time_nodes"]):
            out.append(create_prior_string(f"em_syserr_{gname}_{n}", values.copy()))
    return out


def handle_withoutTime(values):
    validate_distribution(values)
    return [create_prior_string("em_syserr", values)]


config_handlers = {"withTime": handle_withTime, "withoutTime": handle_withoutTime}


def get_prior_strings(yaml_dict):
    validate_only_one_true(yaml_dict)
    out = []
    for key, values in yaml_dict["config"].items():
        if values["value"] and key in config_handlers:
            out.extend(config_handlers[key](values))
    return out


def main(yaml_file_path):
    return get_prior_strings(load_yaml(yaml_file_path))
