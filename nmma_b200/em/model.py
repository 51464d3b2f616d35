"""Light-curve model classes of the hot path.

Mirror of ``nmma/em/model.py``: ``model_parameters_dict`` (``:29-132``),
``LightCurveModelContainer`` (``:175-408``) and ``SVDLightCurveModel`` (``:535-731``).
Same constructor signature, attributes and error behaviour; the light curves themselves
are evaluated by the CUDA engine (``nmma_b200.engine``), never on the CPU.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, Optional, Sequence

import numpy as np

from .. import _lib as L
from .._lib import ParamSrc
from ..core.conversion import (distance_modulus_nmma, get_cosmo_grids, get_redshift,
                               observation_angle_conversion)
from ..mlmodel import SurrogateWeights, load_surrogate, pack_surrogate

# Order matters: it is the input order of the surrogate (nmma/em/model.py:29-132, SVD models only).
model_parameters_dict = {
    "Bu2019nsbh": ["log10_mej_dyn", "log10_mej_wind", "KNtheta"],
    "Bu2019lm": ["log10_mej_dyn", "log10_mej_wind", "KNphi", "KNtheta"],
    "Bu2019lm_sparse": ["log10_mej_dyn", "log10_mej_wind"],
    "Ka2017": ["log10_mej", "log10_vej", "log10_Xlan"],
    "Bu2022mv": ["log10_mej_dyn", "vej_dyn", "log10_mej_wind", "vej_wind", "KNtheta"],
    "Bu2022Ye": ["log10_mej_dyn", "vej_dyn", "Yedyn", "log10_mej_wind", "vej_wind", "KNtheta"],
    "Bu2023Ye": ["log10_mej_dyn", "vej_dyn", "Yedyn", "log10_mej_wind", "vej_wind", "Yewind", "KNtheta"],
    "LANL2022": ["log10_mej_dyn", "vej_dyn", "log10_mej_wind", "vej_wind", "KNtheta"],
    "LANLTP1": ["log10_mej_dyn", "vej_dyn", "log10_mej_wind", "vej_wind", "KNtheta"],
    "LANLTP2": ["log10_mej_dyn", "vej_dyn", "log10_mej_wind", "vej_wind", "KNtheta"],
    "LANLTS1": ["log10_mej_dyn", "vej_dyn", "log10_mej_wind", "vej_wind", "KNtheta"],
    "LANLTS2": ["log10_mej_dyn", "vej_dyn", "log10_mej_wind", "vej_wind", "KNtheta"],
}

citation_dict = {
    **dict.fromkeys(["Bu2019lm", "Bu2019lm_sparse"],
                    ["https://arxiv.org/abs/2002.11355", "https://arxiv.org/abs/1906.04205"]),
    "Bu2019nsbh": ["https://arxiv.org/abs/2009.07210", "https://arxiv.org/abs/1906.04205"],
    **dict.fromkeys(["Bu2022Ye", "Bu2023Ye", "Bu2022mv"],
                    ["https://arxiv.org/abs/2307.11080", "https://arxiv.org/abs/1906.04205"]),
    "Ka2017": ["https://arxiv.org/abs/1710.05463"],
    **dict.fromkeys(["LANLTP1", "LANLTP2", "LANLTS1", "LANLTS2"], ["https://arxiv.org/abs/2105.11543"]),
}


def get_models_home(models_home=None) -> str:
    """``nmma/core/gitlab.py:35-43``: explicit path, else ``$NMMA_MODELS``, else ``./svdmodels``."""
    if models_home is None:
        models_home = os.environ.get("NMMA_MODELS", os.path.join(os.getcwd(), "svdmodels"))
    return os.path.expanduser(models_home)


def resolve_param_sources(model_parameters: Sequence[str], available: Dict[str, ParamSrc]):
    """Device-side version of ``parameter_conversion`` + ``combine_lc_params``
    (``nmma/em/model.py:272-286,701-705``; ``nmma/core/conversion.py:119-126``).

    ``available`` maps every parameter name the sampler provides (column or constant) to its
    source.  Returns the ``ParamSrc`` of each model parameter, in model order.
    """
    avail = dict(available)
    if "KNtheta" not in avail:
        if "inclination_EM" in avail:
            s = avail["inclination_EM"]
            avail["KNtheta"] = ParamSrc(s.col, L.XF_RAD2DEG, s.value)
        elif "theta_jn" in avail:
            s = avail["theta_jn"]
            avail["KNtheta"] = ParamSrc(s.col, L.XF_THETAJN_DEG, s.value)
        elif "cos_theta_jn" in avail:
            s = avail["cos_theta_jn"]
            avail["KNtheta"] = ParamSrc(s.col, L.XF_COSTHETAJN_DEG, s.value)
        else:  # default theta_jn = arccos(1.0) = 0
            avail["KNtheta"] = ParamSrc.const(0.0)
    out = []
    for key in model_parameters:
        if key in avail:
            s = avail[key]
            if s.transform != L.XF_NONE and key != "KNtheta":
                raise ValueError(f"{key}: unexpected transform")
            out.append(s)
        elif key.lstrip("log10_") in avail:          # reference quirk: lstrip strips a character set
            s = avail[key.lstrip("log10_")]
            out.append(ParamSrc(s.col, L.XF_LOG10, s.value))
        elif "log10_" + key in avail:
            s = avail["log10_" + key]
            out.append(ParamSrc(s.col, L.XF_POW10, s.value))
        else:
            # combine_lc_params falls back to getattr(self, key) and raises
            raise AttributeError(f"'SVDLightCurveModel' object has no attribute '{key}'")
    return out


def resolve_constraint_sources(keys: Sequence[str], model_parameters: Sequence[str], available: Dict[str, ParamSrc]):
    """Where the value of each Constraint prior comes from after the reference's conversion chain
    (``nmma/core/base.py:67-68,77-82``): a sampled column / fixed value as is, or a key the chain derives
    (``KNtheta`` from the inclination, ``log10_x`` <-> ``x`` twins of model parameters)."""
    out = []
    for key in keys:
        if key in available:
            out.append(available[key])
        elif key == "KNtheta" or key in model_parameters:
            out.append(resolve_param_sources([key], available)[0])
        else:
            raise NotImplementedError(f"Constraint on '{key}': the EM conversion chain does not produce this key "
                                      "(joint GW-EM constraints are outside the nmma_b200 hot path)")
    return out


class LightCurveModelContainer:
    """Parent class (``nmma/em/model.py:175-408``): detector-frame conversion around ``generate_lightcurve``."""

    extinction_law = "P92_SMC_host"

    def __init__(self, model, filters=None, model_parameters=None, sample_times=None):
        if model_parameters is None:
            assert model in model_parameters_dict.keys(), (
                f"{model} unknown," "please update model_parameters_dict at em/model.py")
            self.model_parameters = model_parameters_dict[model]
        else:
            self.model_parameters = model_parameters
        self.model = model
        self.redshift_func = get_redshift
        self._z_table = None
        if isinstance(filters, str):
            filters = filters.split(",")
        self.filters = filters
        from . import utils
        from ..core.constants import c_SI
        self.default_filts, self.lambdas = utils.get_default_filts_lambdas(self.filters) if self.filters else ([], np.zeros(0))
        self.nu_0s = c_SI / self.lambdas                  # nmma/em/model.py:223-224
        self.good_parameters = True
        self.model_times = sample_times if sample_times is not None else self.setup_model_times()

    def extinction_plan(self, filters):
        """Device staging of ``get_extinction_mags`` (``nmma/em/model.py:323-342``) for ``filters`` (the model filters the
        engine evaluates): ``(law, nu0[F], coef[F])``.  ``nu0 = 0`` marks a filter without a wavelength entry
        (``apply_extinction_correction`` skips it, :344-350)."""
        nu0 = np.array([self.nu_0s[self.default_filts.index(f)] if f in self.default_filts else 0.0 for f in filters], float)
        if self.extinction_law == "P92_SMC_host":
            return L.EXT_P92_SMC_HOST, nu0, None
        if self.extinction_law == "G23_MW":
            # observer-frame, redshift-independent: ext_mag_f = R_V A(lambda_f)/A(V) * Ebv.  The G23 curve itself lives in
            # the third-party dust_extinction package (absent offline); it is evaluated once per filter at staging time
            coef = getattr(self, "extinction_coefficients", None)
            if coef is None:
                try:
                    from dust_extinction.parameter_averages import G23
                    import astropy.units as u
                except ImportError as exc:
                    raise NotImplementedError(
                        "extinction_law 'G23_MW' needs dust_extinction (or model.extinction_coefficients = "
                        "{filter: A_filter / E(B-V)}) to stage the per-filter coefficients") from exc
                law = G23(Rv=3.1)
                x = 1.0 / (self.lambdas * 1e6)
                coef = {f: (3.1 * float(law(xx / u.micron)) if law.x_range[0] <= xx <= law.x_range[1] else 0.0)
                        for f, xx in zip(self.default_filts, x)}
            return L.EXT_LINEAR, None, np.array([float(coef.get(f, 0.0)) for f in filters], float)
        raise ValueError(f"Unknown extinction_law {self.extinction_law!r}use 'P92_SMC_host' or 'G23_MW'.")

    def __repr__(self):
        return self.__class__.__name__ + f"(model={self.model})"

    def setup_model_times(self, tmin=0.01, tmax=14.0, nsteps=150):
        return np.geomspace(tmin, tmax, nsteps)

    def check_vs_priors(self, priors):
        """``:249-267``: build the 50-point dL -> z table when distance (not redshift) is sampled."""
        for key in self.model_parameters:
            if key not in priors:
                print(f"Parameter {key} not found in priors, might fail.", file=sys.stderr)   # reference prints this to stdout
        if "redshift" not in priors and "luminosity_distance" in priors:
            dl = priors["luminosity_distance"]
            lo = getattr(dl, "minimum", None)
            hi = getattr(dl, "maximum", None)
            if lo is None:                      # fixed distance given as a bare number
                lo = hi = float(dl)
            dist_grid, z_grid = get_cosmo_grids(lo, hi)
            self._z_table = (np.asarray(dist_grid, float), np.asarray(z_grid, float))

            def redshift_from_dlum(parameters):
                return np.interp(parameters["luminosity_distance"], dist_grid, z_grid)

            self.redshift_func = redshift_from_dlum

    def sanity_checks(self, parameters):
        self.good_parameters = True

    def parameter_conversion(self, parameters):
        """``:272-286`` (host mirror; the batched path does the same on the device)."""
        new = observation_angle_conversion(parameters)
        for key in self.model_parameters:
            if key not in new:
                if key.lstrip("log10_") in new.keys():
                    new[key] = np.log10(new[key.lstrip("log10_")])
                elif "log10_" + key in new.keys():
                    new[key] = 10 ** new["log10_" + key]
        self.sanity_checks(new)
        return new

    def em_parameter_setup(self, parameters, combine_params=True):
        """``:288-303``."""
        self.Ebv = parameters.get("Ebv", 0.0)
        self.luminosity_distance = parameters.get("luminosity_distance", 1e-5)
        self.distmod = distance_modulus_nmma(self.luminosity_distance)
        self.timeshift = parameters.get("timeshift", 0.0)
        self.redshift = self.redshift_func(parameters)
        if combine_params:
            return self.combine_lc_params(parameters)

    def combine_lc_params(self, parameters):
        return {k: parameters[k] if k in parameters else getattr(self, k) for k in self.model_parameters}

    def generate_lightcurve(self, sample_times, parameters):
        raise NotImplementedError("This method should be implemented in subclasses.")

    @property
    def citation(self):
        return {self.model: citation_dict[self.model]}


class SVDLightCurveModel(LightCurveModelContainer):
    """SVD-surrogate light curves evaluated on the GPU (``nmma/em/model.py:535-731``).

    Parameters are those of the reference.  ``interpolation_type`` accepts ``tensorflow`` /
    ``keras`` (per-filter MLP) and ``sklearn_gp``; ``api_gp`` is out of scope.
    ``svd_mag_model`` may be passed directly (reference in-memory layout with per-filter
    ``'model'`` = (W1, b1, W2, b2) or ``'gps'``) to skip file loading, e.g. for random-init weights.
    """

    def __init__(self, model, svd_path=None, svd_mag_ncoeff=None, svd_lbol_ncoeff=None,
                 interpolation_type="keras", model_parameters=None, filters=None, sample_times=None,
                 local_only=False, svd_mag_model=None, device=0, extinction_law=None, **em_model_kwargs):
        comps = model.split("_")
        if "tf" in comps:
            comps.remove("tf")
        core_model_name = "_".join(comps)
        self.mag_ncoeff = svd_mag_ncoeff
        self.lbol_ncoeff = svd_lbol_ncoeff
        self.interpolation_type = interpolation_type
        self.svd_path = get_models_home(svd_path)
        self.model_specifier = "_tf" if interpolation_type == "tensorflow" else ""
        self.device = device
        if isinstance(filters, str):
            filters = filters.split(",")
        if interpolation_type not in ("sklearn_gp", "keras", "tensorflow", "torch", "jax"):
            if interpolation_type == "api_gp":
                raise ValueError("--interpolation-type api_gp is not supported by nmma_b200")
            raise ValueError("--interpolation-type must be sklearn_gp, api_gp or tensorflow")
        if svd_mag_model is None:
            # no network in this stack: behaves like --local-only (nmma/core/gitlab.py is out of scope)
            core, filters, found, kind = load_surrogate(model, self.svd_path, filters, interpolation_type,
                                                        svd_mag_ncoeff)
        else:
            core = svd_mag_model
            kind = "gp" if interpolation_type == "sklearn_gp" else "mlp"
            if filters is None:
                filters = list(core.keys())
            found = [f for f in filters if f in core and ("gps" if kind == "gp" else "model") in core[f]]
        self.svd_mag_model = core
        self.svd_lbol_model = None
        self._kind = kind
        super().__init__(core_model_name, filters, model_parameters, sample_times)
        # filters the surrogate can evaluate; the rest produce +inf light curves (lightcurve_generation.py:171)
        self._eval_filters = [f for f in self.filters if f in found]
        if not self._eval_filters:
            raise ValueError(f"No model files found for {model}")
        self.weights: SurrogateWeights = pack_surrogate(core, self._eval_filters, kind, svd_mag_ncoeff)
        self._engine = None
        if extinction_law is not None:        # create_light_curve_model_from_args, nmma/em/model.py:1611-1613
            self.extinction_law = extinction_law

    # ---- engine plumbing ----------------------------------------------------------------
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_engine"] = None          # device handles never travel through pickles
        state["redshift_func"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        if self._z_table is not None:
            dg, zg = self._z_table
            self.redshift_func = lambda p: np.interp(p["luminosity_distance"], dg, zg)
        else:
            self.redshift_func = get_redshift

    def new_engine(self):
        """A fresh engine with this model's surrogate staged (used by the likelihood)."""
        from ..engine import KilonovaEngine
        eng = KilonovaEngine(self.device)
        eng.set_surrogate(self.weights)
        return eng

    def _canonical_engine(self, sample_times):
        """Engine whose points are [x_0..x_{d-1}, luminosity_distance, timeshift, redshift, Ebv]."""
        if self._engine is None or getattr(self, "_engine_law", None) != self.extinction_law:
            eng = self.new_engine()
            d = self.weights.d
            eng.set_param_layout(d + 4, [ParamSrc.column(i) for i in range(d)],
                                 ParamSrc.column(d), ParamSrc.column(d + 1), ParamSrc.column(d + 2), L.Z_PARAM)
            law, nu0, coef = self.extinction_plan(self._eval_filters)
            eng.set_extinction(law, ParamSrc.column(d + 3), nu0, coef)
            self._engine = eng
            self._engine_law = self.extinction_law
            self._engine_grid = None
        st = np.ascontiguousarray(sample_times, float)
        if self._engine_grid is None or self._engine_grid.shape != st.shape or not np.array_equal(self._engine_grid, st):
            self._engine.set_sample_grid(st)
            self._engine_grid = st.copy()
        return self._engine

    # ---- reference API ------------------------------------------------------------------
    def setup_model_times(self):
        try:
            return next(iter(self.svd_mag_model.values()))["tt"]
        except Exception:
            return super().setup_model_times()

    def __repr__(self):
        return super().__repr__() + f"(model={self.model}, svd_path={self.svd_path})"

    def combine_lc_params(self, parameters):
        return [parameters[k] if k in parameters else getattr(self, k) for k in self.model_parameters]

    def _row(self, parameters):
        plist = self.em_parameter_setup(parameters)
        return np.array([list(plist) + [self.luminosity_distance, self.timeshift, self.redshift, self.Ebv]], float)

    def generate_lightcurve(self, sample_times, parameters, filters="all"):
        """Absolute AB magnitudes on ``sample_times`` per filter (``:707-728``), evaluated on the GPU."""
        if filters is None:
            raise NotImplementedError("bolometric SVD models are not implemented upstream either (svd_lbol_model is None)")
        if filters == "all":
            filters = self.filters
        sample_times = np.asarray(sample_times, float)
        eng = self._canonical_engine(sample_times)
        mags, _ = eng.mags(self._row(parameters), apparent=False)
        mags = mags[0].cpu().numpy()
        out = {}
        for filt in filters:
            if filt in self._eval_filters:
                out[filt] = mags[self._eval_filters.index(filt)].copy()
            else:
                out[filt] = np.full_like(sample_times, np.inf)
        return out

    def gen_detector_lc(self, parameters=None, sample_times=None):
        """(observable_times, {filt: apparent mags}) in the detector frame (``:352-404``)."""
        if sample_times is None:
            sample_times = self.model_times
        sample_times = np.asarray(sample_times, float)
        eng = self._canonical_engine(sample_times)
        mags, tobs = eng.mags(self._row(parameters), apparent=True)
        mags = mags[0].cpu().numpy()
        tobs = tobs[0].cpu().numpy()
        lc = {}
        for filt in self.filters:
            if filt in self._eval_filters:
                lc[filt] = mags[self._eval_filters.index(filt)].copy()
            else:
                lc[filt] = np.full_like(tobs, np.inf)
        return tobs, lc

    def generate_spectra(self, sample_times, wavelengths, parameters):
        return self.generate_lightcurve(sample_times, parameters, filters=wavelengths)


class CombinedLightCurveModelContainer:
    """``nmma/em/model.py:1342-1510``: several light-curve models evaluated on the same parameters, their light curves added
    in flux (``stack_magnitudes``: -2.5 log10 sum 10^(-0.4 m), through logsumexp like the reference).

    Host-side composition: each sub-model (an :class:`SVDLightCurveModel`) evaluates its magnitudes on the GPU, the stack is
    a handful of vector operations per call.  That serves ``generate_lightcurve`` / ``gen_detector_lc`` (best-fit plots,
    injections).  The batched likelihood needs the flux sum on the device before the interpolation to the observation
    times and is not implemented for combined models: ``EMTransientLikelihood`` raises for them."""

    def __init__(self, models, model_args=None):
        from . import utils
        self.lc_models = list(models) if model_args is None else [m(*model_args[i]) for i, m in enumerate(models)]
        self.model = [m.model for m in self.lc_models]
        self.all_filters = set().union(*[m.filters for m in self.lc_models])
        self.filters = sorted(self.all_filters)
        self.compatible_filters, _ = utils.get_filter_name_mapping(self.all_filters, extra_known=tuple(self.all_filters))
        self.model_times = np.array(sorted(set().union(*[np.asarray(m.model_times, float).tolist() for m in self.lc_models])))
        self.model_parameters = [k for m in self.lc_models for k in m.model_parameters]

    def __repr__(self):
        return "Combination of " + " and ".join(repr(m) for m in self.lc_models)

    def check_vs_priors(self, priors):
        for m in self.lc_models:
            m.check_vs_priors(priors)

    @property
    def citation(self):
        out = {}
        for m in self.lc_models:
            out.update(m.citation)
        return out

    @property
    def good_parameters(self):
        return bool(np.prod([getattr(m, "good_parameters", True) for m in self.lc_models]))

    def parameter_conversion(self, parameters):
        for m in self.lc_models:
            parameters = m.parameter_conversion(parameters)
        return parameters

    def new_engine(self):
        raise NotImplementedError("combined light-curve models have no device engine: the batched likelihood covers single "
                                  "SVD surrogates (generate_lightcurve / gen_detector_lc of the combination are available)")

    def stack_magnitudes(self, mags_per_model):
        """``:1486-1510``: per filter, the flux sum of the models that provide it (directly or as a filter average)."""
        from scipy.special import logsumexp
        from . import utils
        ln10 = np.log(10.0)
        stacked = {}
        for filt in self.all_filters:
            terms = []
            for mag in mags_per_model:
                try:
                    m = mag[self.compatible_filters[filt]]
                except KeyError:
                    try:
                        m = utils.average_mags(mag, filt)
                    except (ValueError, KeyError):
                        continue
                terms.append(-2.0 / 5.0 * ln10 * np.asarray(m, float))
            if not terms:
                stacked[filt] = np.full_like(self.model_times, np.inf)
            else:
                stacked[filt] = -5.0 / 2.0 * logsumexp(terms, axis=0) / ln10
        return stacked

    def generate_lightcurve(self, sample_times, parameters, return_all=False):
        per_model = []
        for m in self.lc_models:
            lc = m.generate_lightcurve(sample_times, parameters)
            if not lc:
                return lc
            per_model.append(lc)
        return per_model if return_all else self.stack_magnitudes(per_model)

    def gen_detector_lc(self, parameters, sample_times=None, return_all=False):
        """``:1410-1460``: every model in its own detector frame, brought onto the union of the time grids
        (``autocomplete_data`` with +inf outside a model's range), then stacked."""
        from . import utils
        lcs, times = [], []
        for m in self.lc_models:
            t, lc = m.gen_detector_lc(parameters, sample_times)
            if not lc:
                return t, lc
            lcs.append(lc)
            times.append(np.asarray(t, float))
        if return_all:
            return times, lcs
        joint = np.array(sorted(set().union(*[t.tolist() for t in times]))) if sample_times is None else times[-1]
        on_joint = [{f: utils.autocomplete_data(joint, t, v, extrapolate=np.inf) for f, v in lc.items()}
                    for t, lc in zip(times, lcs)]
        return joint, self.stack_magnitudes(on_joint)


GenericCombineLightCurveModel = CombinedLightCurveModelContainer   # the reference's legacy synonym (``:1513``)


def create_light_curve_model_from_args(model_name_arg, args, filters=None, sample_times=None):
    """``nmma/em/model.py:1617-1658`` restricted to SVD kilonova models."""
    from .utils import setup_sample_times
    if filters is None:
        from .utils import set_filters
        filters = set_filters(args)
    if sample_times is None:
        sample_times = setup_sample_times(args)
    names = model_name_arg.split(",") if isinstance(model_name_arg, str) else list(model_name_arg)
    def one(name):
        return SVDLightCurveModel(
            name, svd_path=getattr(args, "svd_path", None),
            extinction_law=getattr(args, "em_extinction_law", None),
            svd_mag_ncoeff=getattr(args, "svd_mag_ncoeff", None),
            svd_lbol_ncoeff=getattr(args, "svd_lbol_ncoeff", None),
            interpolation_type=getattr(args, "interpolation_type", "keras"),
            filters=filters, sample_times=sample_times, local_only=getattr(args, "local_only", True))

    if len(names) != 1:       # SVD surrogates only (afterglowpy / analytic models are outside the path)
        return CombinedLightCurveModelContainer([one(n) for n in names])
    return one(names[0])
