"""Likelihood base classes: the drop-in boundary towards bilby samplers.

Mirror of ``nmma/core/base.py:37-185`` (``NMMALikelihoodMixin`` / ``NMMALikelihood``):
conversion chain -> constraint product -> sub-model log-likelihood, with the
reference's failure sentinel ``np.nan_to_num(-np.inf)`` instead of exceptions.  bilby is
optional: when it is importable its ``Likelihood`` is the base class so the object drops
into ``bilby.run_sampler`` unchanged; otherwise a minimal stand-in with the same surface
(``parameters``, ``log_likelihood``, ``noise_log_likelihood``, ``log_likelihood_ratio``)
is used.
"""
from __future__ import annotations

import inspect

import numpy as np

from .constants import SENTINEL
from .priors import is_constraint

try:  # pragma: no cover - bilby is not installed in the offline image
    from bilby.core.likelihood import Likelihood as _BilbyLikelihood
except Exception:  # noqa: BLE001
    _BilbyLikelihood = None


class Likelihood(object if _BilbyLikelihood is None else _BilbyLikelihood):
    """Stand-in for ``bilby.core.likelihood.Likelihood`` (same public surface)."""

    if _BilbyLikelihood is None:
        def __init__(self, parameters=None):
            self.parameters = parameters if parameters is not None else {}
            self._meta_data = None
            self._marginalized_parameters = []

        def log_likelihood(self, parameters=None):
            raise NotImplementedError

        def noise_log_likelihood(self):
            return np.nan

        def log_likelihood_ratio(self, parameters=None):
            return self.log_likelihood(parameters) - self.noise_log_likelihood()

        @property
        def meta_data(self):
            return self._meta_data

        @meta_data.setter
        def meta_data(self, value):
            self._meta_data = value

        @property
        def marginalized_parameters(self):
            return self._marginalized_parameters


def initialisation_args_from_signature_and_namespace(_callable, namespace, prefixes=None):
    """``nmma/core/base.py:20-35``: map CLI names (with ``em_``/``kilonova_`` prefixes) onto kwargs."""
    prefixes = list(prefixes or []) + [""]
    signature = inspect.signature(_callable)
    kwargs = {k: v.default for k, v in signature.parameters.items() if v.default is not inspect.Parameter.empty}
    for key in signature.parameters.keys():
        for prefix in prefixes:
            if hasattr(namespace, prefix + key):
                val = getattr(namespace, prefix + key)
                if val is not None:
                    kwargs[key] = val
                break
    return kwargs


class NMMALikelihoodMixin:
    """``nmma/core/base.py:37-132``."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    @property
    def priors(self):
        return self._priors

    @priors.setter
    def priors(self, value):
        self.constraints = value
        sampling_keys = [k for k in value.keys() if k not in self.constraints]
        self.check_parameter_equivalencies(sampling_keys)
        self._priors = value

    @property
    def constraints(self):
        return self._constraints

    @constraints.setter
    def constraints(self, value):
        if is_constraint(value):
            constr = {value.name: value}
        elif hasattr(value, "items"):
            constr = {k: v for k, v in value.items() if is_constraint(v)}
        else:
            constr = {}
        self._constraints = constr

    def evaluate_constraints(self, out_sample):
        return np.prod([con.prob(out_sample[k]) for k, con in self.constraints.items()])

    def identity_conversion(self, parameters):
        return parameters

    def __call__(self, parameters):
        return np.exp(self.log_likelihood(parameters))

    def log_likelihood(self, parameters):
        parameters = self.parameter_conversion(parameters)
        if self.evaluate_constraints(parameters) and self.sanity_checks():
            return self.sub_log_likelihood(parameters)
        return SENTINEL

    def sanity_checks(self):
        return True

    def check_parameter_equivalencies(self, parameter_names):
        """``nmma/core/base.py:111-131``."""
        for group in (["inclination_EM", "KNtheta", "theta_jn", "cos_theta_jn", "thetaObs"],):
            inter = set(parameter_names).intersection(group)
            if len(inter) > 1:
                raise ValueError(f"Multiple equivalent parameters found: {inter}. Please only provide one of these.")
        for group in (["redshift", "luminosity_distance", "Hubble_constant"],
                      ["mass_1", "mass_1_source", "chirp_mass", "mass_ratio", "eta", "mass_2", "mass_2_source"]):
            inter = set(parameter_names).intersection(group)
            if len(inter) > 2:
                raise ValueError(f"Mutually dependent parameters found: {inter}. Please only provide up to two of these.")


class NMMALikelihood(NMMALikelihoodMixin, Likelihood):
    """``nmma/core/base.py:134-185``."""

    def __init__(self, sub_model, priors, **kwargs):
        super().__init__()
        self.sub_model = sub_model
        try:
            self._noise_logl = self.sub_model.noise_log_likelihood()
        except AttributeError:
            self._noise_logl = 0.0
        self.conv_functions = []
        self.priors = priors
        self.setup_submodel_conversion()

    def __repr__(self):
        return self.__class__.__name__ + " with " + self.sub_model.__repr__()

    def setup_submodel_conversion(self):
        pass

    def parameter_conversion(self, parameters):
        for conv in reversed(self.conv_functions):
            parameters = conv(parameters)
        return parameters

    def posterior_conversion(self, parameters):
        return self.parameter_conversion(parameters)

    def sub_log_likelihood(self, parameters):
        logl = self.sub_model.log_likelihood(parameters)
        if not np.isfinite(logl):
            return SENTINEL
        return logl

    def noise_log_likelihood(self):
        return self._noise_logl
