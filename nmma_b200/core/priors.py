"""Minimal prior layer for the kilonova configurations (bilby is absent offline).

The reference builds its priors with ``bilby.core.prior.PriorDict(prior_file)``
(``nmma/em/prior.py:221-244``).  This module restates the subset of bilby's prior
classes and prior-file grammar the shipped kilonova priors use (SURVEY.md
Appendix C): ``key = Class(kw=...)`` lines or bare float literals, ``np.pi``
arithmetic allowed, the LHS key is the parameter name.  The objects expose the
bilby attributes the hot path touches (``minimum``, ``maximum``, ``name``,
``rescale``, ``sample``, ``prob``/``ln_prob``, ``peak`` for fixed values), so a
genuine bilby ``PriorDict`` can be passed to the likelihood interchangeably.
"""
from __future__ import annotations

import re
from collections import OrderedDict

import numpy as np
from scipy.special import erf, erfinv


class Prior:
    _args = ("name", "latex_label", "unit", "boundary")

    def __init__(self, name=None, latex_label=None, unit=None, minimum=-np.inf, maximum=np.inf, boundary=None, **_ignored):
        self.name = name
        self._latex_label = latex_label
        self.unit = unit
        self.minimum = minimum
        self.maximum = maximum
        self.boundary = boundary

    @property
    def latex_label(self):
        return self._latex_label if self._latex_label is not None else self.name

    @latex_label.setter
    def latex_label(self, v):
        self._latex_label = v

    @property
    def is_fixed(self):
        return False

    def rescale(self, val):
        raise NotImplementedError

    def prob(self, val):
        raise NotImplementedError

    def ln_prob(self, val):
        with np.errstate(divide="ignore"):
            return np.log(self.prob(val))

    def sample(self, size=None, rng=None):
        rng = rng if rng is not None else np.random.default_rng()
        return self.rescale(rng.uniform(0, 1, size))

    def is_in_prior_range(self, val):
        return (val >= self.minimum) & (val <= self.maximum)

    def __repr__(self):
        args = ", ".join(f"{k}={getattr(self, k)!r}" for k in self._args)
        return f"{self.__class__.__name__}({args})"


class Uniform(Prior):
    _args = ("minimum", "maximum", "name", "latex_label", "unit", "boundary")

    def __init__(self, minimum, maximum, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name, latex_label, unit, minimum, maximum, boundary)

    def rescale(self, val):
        return self.minimum + val * (self.maximum - self.minimum)

    def prob(self, val):
        return ((val >= self.minimum) & (val <= self.maximum)) / (self.maximum - self.minimum)


class DeltaFunction(Prior):
    _args = ("peak", "name", "latex_label", "unit")

    def __init__(self, peak, name=None, latex_label=None, unit=None):
        super().__init__(name, latex_label, unit, peak, peak)
        self.peak = peak

    @property
    def is_fixed(self):
        return True

    def rescale(self, val):
        return self.peak * val ** 0

    def prob(self, val):
        at_peak = (val == self.peak)
        return np.nan_to_num(np.multiply(at_peak, np.inf))


class Sine(Prior):
    _args = ("minimum", "maximum", "name", "latex_label", "unit", "boundary")

    def __init__(self, minimum=0, maximum=np.pi, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name, latex_label, unit, minimum, maximum, boundary)

    def rescale(self, val):
        norm = 1 / (np.cos(self.minimum) - np.cos(self.maximum))
        return np.arccos(np.cos(self.minimum) - val / norm)

    def prob(self, val):
        return np.sin(val) / 2 * self.is_in_prior_range(val)


class Cosine(Prior):
    _args = ("minimum", "maximum", "name", "latex_label", "unit", "boundary")

    def __init__(self, minimum=-np.pi / 2, maximum=np.pi / 2, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name, latex_label, unit, minimum, maximum, boundary)

    def rescale(self, val):
        norm = 1 / (np.sin(self.maximum) - np.sin(self.minimum))
        return np.arcsin(val / norm + np.sin(self.minimum))

    def prob(self, val):
        return np.cos(val) / 2 * self.is_in_prior_range(val)


class Gaussian(Prior):
    _args = ("mu", "sigma", "name", "latex_label", "unit", "boundary")

    def __init__(self, mu, sigma, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name, latex_label, unit, -np.inf, np.inf, boundary)
        self.mu, self.sigma = mu, sigma

    def rescale(self, val):
        return self.mu + erfinv(2 * val - 1) * 2 ** 0.5 * self.sigma

    def prob(self, val):
        return np.exp(-(self.mu - val) ** 2 / (2 * self.sigma ** 2)) / (2 * np.pi) ** 0.5 / self.sigma


Normal = Gaussian


class TruncatedGaussian(Prior):
    _args = ("mu", "sigma", "minimum", "maximum", "name", "latex_label", "unit", "boundary")

    def __init__(self, mu, sigma, minimum, maximum, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name, latex_label, unit, minimum, maximum, boundary)
        self.mu, self.sigma = mu, sigma

    @property
    def normalisation(self):
        return (erf((self.maximum - self.mu) / 2 ** 0.5 / self.sigma)
                - erf((self.minimum - self.mu) / 2 ** 0.5 / self.sigma)) / 2

    def rescale(self, val):
        return erfinv(2 * val * self.normalisation
                      + erf((self.minimum - self.mu) / 2 ** 0.5 / self.sigma)) * 2 ** 0.5 * self.sigma + self.mu

    def prob(self, val):
        return np.exp(-(self.mu - val) ** 2 / (2 * self.sigma ** 2)) / (2 * np.pi) ** 0.5 \
            / self.sigma / self.normalisation * self.is_in_prior_range(val)


TruncatedNormal = TruncatedGaussian


class PowerLaw(Prior):
    _args = ("alpha", "minimum", "maximum", "name", "latex_label", "unit", "boundary")

    def __init__(self, alpha, minimum, maximum, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(name, latex_label, unit, minimum, maximum, boundary)
        self.alpha = alpha

    def rescale(self, val):
        if self.alpha == -1:
            return self.minimum * np.exp(val * np.log(self.maximum / self.minimum))
        a1 = 1 + self.alpha
        return (self.minimum ** a1 + val * (self.maximum ** a1 - self.minimum ** a1)) ** (1.0 / a1)

    def prob(self, val):
        if self.alpha == -1:
            return np.nan_to_num(1 / val / np.log(self.maximum / self.minimum)) * self.is_in_prior_range(val)
        a1 = 1 + self.alpha
        return np.nan_to_num(val ** self.alpha * a1 / (self.maximum ** a1 - self.minimum ** a1)) \
            * self.is_in_prior_range(val)


class LogUniform(PowerLaw):
    _args = ("minimum", "maximum", "name", "latex_label", "unit", "boundary")

    def __init__(self, minimum, maximum, name=None, latex_label=None, unit=None, boundary=None):
        super().__init__(-1, minimum, maximum, name, latex_label, unit, boundary)


class Triangular(Prior):
    _args = ("mode", "minimum", "maximum", "name", "latex_label", "unit")

    def __init__(self, mode, minimum, maximum, name=None, latex_label=None, unit=None):
        super().__init__(name, latex_label, unit, minimum, maximum)
        self.mode = mode

    def rescale(self, val):
        a, b, c = self.minimum, self.maximum, self.mode
        fc = (c - a) / (b - a)
        val = np.asarray(val, float)
        lo = a + np.sqrt(np.maximum(val, 0) * (b - a) * (c - a))
        hi = b - np.sqrt(np.maximum(1 - val, 0) * (b - a) * (b - c))
        return np.where(val < fc, lo, hi)

    def prob(self, val):
        a, b, c = self.minimum, self.maximum, self.mode
        val = np.asarray(val, float)
        with np.errstate(divide="ignore", invalid="ignore"):
            up = 2 * (val - a) / ((b - a) * (c - a))
            dn = 2 * (b - val) / ((b - a) * (b - c))
        return np.where((val >= a) & (val <= c) & (c > a), up, np.where((val > c) & (val <= b), dn, 0.0)) \
            if c > a else np.where((val >= a) & (val <= b), dn, 0.0)


class Interped(Prior):
    """Tabulated density (the reference's triangular Ebv prior, ``nmma/em/prior.py:209-216``)."""
    _args = ("xx", "yy", "minimum", "maximum", "name", "latex_label", "unit", "boundary")

    def __init__(self, xx, yy, minimum=np.nan, maximum=np.nan, name=None, latex_label=None, unit=None, boundary=None):
        self.xx = np.asarray(xx, float)
        self.yy = np.asarray(yy, float)
        minimum = np.nanmax([np.min(self.xx), minimum])
        maximum = np.nanmin([np.max(self.xx), maximum])
        super().__init__(name, latex_label, unit, minimum, maximum, boundary)
        grid = np.linspace(self.minimum, self.maximum, 2001)
        pdf = np.interp(grid, self.xx, self.yy)
        cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(grid))])
        self._grid, self._pdf, self._cdf = grid, pdf / cdf[-1], cdf / cdf[-1]

    def rescale(self, val):
        return np.interp(val, self._cdf, self._grid)

    def prob(self, val):
        return np.interp(val, self._grid, self._pdf, left=0.0, right=0.0)


class Constraint(Prior):
    _args = ("minimum", "maximum", "name", "latex_label", "unit")

    def __init__(self, minimum, maximum, name=None, latex_label=None, unit=None):
        super().__init__(name, latex_label, unit, minimum, maximum)

    def prob(self, val):
        return (val > self.minimum) & (val < self.maximum)


_CLASSES = {c.__name__: c for c in (Uniform, DeltaFunction, Sine, Cosine, Gaussian, TruncatedGaussian, PowerLaw,
                                    LogUniform, Triangular, Interped, Constraint)}
_CLASSES.update(Normal=Gaussian, TruncatedNormal=TruncatedGaussian)


def is_fixed_prior(prior) -> bool:
    """True for DeltaFunction-like priors and bare numbers (ours or bilby's)."""
    if isinstance(prior, (int, float, np.floating, np.integer)):
        return True
    if getattr(prior, "is_fixed", False):
        return True
    return hasattr(prior, "peak") and not hasattr(prior, "rescale_table")


def fixed_value(prior) -> float:
    if isinstance(prior, (int, float, np.floating, np.integer)):
        return float(prior)
    return float(prior.peak)


def is_constraint(prior) -> bool:
    return prior.__class__.__name__ == "Constraint"


def prior_from_string(expr: str, name: str = None):
    """Evaluate the right-hand side of a prior-file line."""
    expr = expr.strip()
    ns = {"np": np, "numpy": np, "__builtins__": {}}
    ns.update(_CLASSES)
    m = re.match(r"^([A-Za-z_][\w\.]*)\s*\(", expr)
    if m and m.group(1).split(".")[-1] not in _CLASSES and not m.group(1).startswith("np."):
        raise ValueError(f"prior class {m.group(1)} is outside the kilonova subset supported by nmma_b200")
    if m and "." in m.group(1) and not m.group(1).startswith("np."):
        expr = m.group(1).split(".")[-1] + expr[m.end(1):]
    val = eval(expr, ns)  # noqa: S307 - prior files are trusted user configuration, as in bilby
    if isinstance(val, Prior):
        if name is not None:
            val.name = name
        return val
    return DeltaFunction(float(val), name=name)


class PriorDict(OrderedDict):
    """Ordered name -> prior mapping with bilby's ``PriorDict`` constructor and sampling surface."""

    def __init__(self, dictionary=None, filename=None):
        super().__init__()
        if isinstance(dictionary, str) and filename is None:
            filename, dictionary = dictionary, None
        if filename is not None:
            self.from_file(filename)
        elif dictionary is not None:
            for k, v in dictionary.items():
                self[k] = prior_from_string(v, k) if isinstance(v, str) else v

    def from_file(self, filename):
        with open(filename) as fh:
            for line in fh:
                line = line.split("#")[0].strip() if not re.search(r"['\"].*#.*['\"]", line) else line.strip()
                if not line or line.startswith("#"):
                    continue
                key, _, rhs = line.partition("=")
                self[key.strip()] = prior_from_string(rhs, key.strip())

    @property
    def fixed_keys(self):
        return [k for k, p in self.items() if is_fixed_prior(p) and not is_constraint(p)]

    @property
    def constraint_keys(self):
        return [k for k, p in self.items() if is_constraint(p)]

    @property
    def non_fixed_keys(self):
        return [k for k, p in self.items() if not is_fixed_prior(p) and not is_constraint(p)]

    def sample(self, size=None, rng=None):
        rng = rng if rng is not None else np.random.default_rng()
        return {k: (self[k].sample(size, rng) if not is_fixed_prior(self[k]) else
                    (fixed_value(self[k]) if size is None else np.full(size, fixed_value(self[k]))))
                for k in self if not is_constraint(self[k])}

    def rescale(self, keys, theta):
        theta = np.asarray(theta, float)
        return [self[k].rescale(theta[..., i]) for i, k in enumerate(keys)]

    def sample_array(self, n, rng=None, keys=None):
        """points[n, P] for the sampled (non-fixed) keys, column order = prior order."""
        keys = list(keys) if keys is not None else self.non_fixed_keys
        rng = rng if rng is not None else np.random.default_rng()
        u = rng.uniform(0, 1, size=(n, len(keys)))
        return np.stack([np.asarray(self[k].rescale(u[:, i]), float) for i, k in enumerate(keys)], axis=1), keys

    def device_plan(self, keys=None):
        """``(kinds[P], params[P,4], tables{j: (cdf, grid)})`` for ``nmma_b200_set_priors``: one analytic
        prior per column of ``points[N,P]`` (``keys`` default: the sampled keys in prior order)."""
        from .. import _lib as L
        keys = list(keys) if keys is not None else self.non_fixed_keys
        kinds, params, tables = [], np.zeros((len(keys), 4)), {}
        for j, k in enumerate(keys):
            p = self[k]
            if is_fixed_prior(p):
                kinds.append(L.PR_DELTA); params[j, 0] = fixed_value(p)
            elif isinstance(p, Uniform):
                kinds.append(L.PR_UNIFORM); params[j, :2] = p.minimum, p.maximum
            elif isinstance(p, Sine):
                kinds.append(L.PR_SINE); params[j, :2] = p.minimum, p.maximum
            elif isinstance(p, Cosine):
                kinds.append(L.PR_COSINE); params[j, :2] = p.minimum, p.maximum
            elif isinstance(p, TruncatedGaussian):
                kinds.append(L.PR_TRUNC_GAUSS); params[j] = p.mu, p.sigma, p.minimum, p.maximum
            elif isinstance(p, Gaussian):
                kinds.append(L.PR_GAUSSIAN); params[j, :2] = p.mu, p.sigma
            elif isinstance(p, PowerLaw):
                kinds.append(L.PR_POWERLAW); params[j, :3] = p.alpha, p.minimum, p.maximum
            elif isinstance(p, Triangular):
                kinds.append(L.PR_TRIANGULAR); params[j, :3] = p.mode, p.minimum, p.maximum
            elif isinstance(p, Interped):
                kinds.append(L.PR_INTERPED); tables[j] = (p._cdf, p._grid)
            else:
                raise NotImplementedError(f"prior {k} = {p!r} has no device transform")
        return np.asarray(kinds, np.int32), params, tables

    def ln_prob(self, sample):
        return float(np.sum([self[k].ln_prob(sample[k]) for k in sample if k in self and not is_constraint(self[k])]))
