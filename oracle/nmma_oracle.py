"""TEST INFRASTRUCTURE ONLY -- never imported by the product (nmma_b200/).

CPU oracle: a NumPy / SciPy / scikit-learn restatement of NMMA's inner
kilonova likelihood loop, one parameter point per call exactly as bilby drives
the reference.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.

Every function cites the reference lines it follows (paths relative to the
reference checkout, nmma v1.0.1 @ 5771cfdb).  The same third-party calls are
used where they are installed (``np.interp``, ``np.dot``,
``scipy.stats.truncnorm.logpdf``, ``scipy.stats.norm.logsf``,
``GaussianProcessRegressor.predict``).  Keras/TensorFlow are not installable
here; the per-filter network ``Input(d) -> Dense(2048, relu) -> Dropout ->
Dense(n_coeff)`` (``nmma/em/training.py:353-364``; dropout is the identity at
inference) is evaluated by :class:`KerasStandIn` in float32 like Keras does.

Parity pinning: the magnitude path is pinned by the reference's
only known-answer test through the SVD surrogate,
``nmma/tests/joint_analysis_pipeline.py:108-120`` (tests/test_oracle_golden.py).
No reference test asserts a log-likelihood value; that part is pinned by running
the reference's *own source files* in this container with the missing third-party
imports stubbed (tests/golden/make_reference_vectors.py).
"""
from __future__ import annotations

import numpy as np
from scipy.stats import norm, truncnorm

from . import cosmology as _cosmo
from . import extinction as _ext

SENTINEL = float(np.nan_to_num(-np.inf))  # -1.7976931348623157e308, nmma/core/base.py:82


# --------------------------------------------------------------------------
# surrogate evaluation
# --------------------------------------------------------------------------
class _Tensor:
    def __init__(self, a):
        self._a = a

    def numpy(self):
        return self._a


class KerasStandIn:
    """float32 stand-in for the Keras model object stored under ``svd_model[filt]['model']``.

    Call convention follows ``nmma/em/lightcurve_generation.py:198``:
    ``model(np.atleast_2d(x)).numpy().T.flatten()``.  Keras casts the float64
    input to float32 and runs both Dense layers in float32.
    """

    def __init__(self, W1, b1, W2, b2):
        self.W1 = np.asarray(W1, np.float32)
        self.b1 = np.asarray(b1, np.float32)
        self.W2 = np.asarray(W2, np.float32)
        self.b2 = np.asarray(b2, np.float32)

    def __call__(self, x):
        x = np.asarray(x).astype(np.float32)
        h = np.maximum(x @ self.W1 + self.b1, np.float32(0))
        return _Tensor(h @ self.W2 + self.b2)


def autocomplete_data(interp_points, ref_points, ref_data, extrapolate="linear", ref_value=np.inf):
    """``nmma/em/utils.py:626-677`` (spline branch omitted: never taken on this path)."""
    data_mask = np.isfinite(ref_data)
    if np.sum(data_mask) < 2:
        return np.full_like(interp_points, ref_value)
    fin_ref = np.asarray(ref_points)[data_mask]
    fin_data = np.asarray(ref_data)[data_mask]
    interp_points = np.atleast_1d(interp_points)
    if isinstance(extrapolate, (float, int)):
        return np.interp(interp_points, fin_ref, fin_data, left=extrapolate, right=extrapolate)
    if isinstance(extrapolate, str):
        if extrapolate == "linear":
            out = np.interp(interp_points, fin_ref, fin_data)
            x0, x1, xm, xn = fin_ref[[0, 1, -2, -1]]
            y0, y1, ym, yn = fin_data[[0, 1, -2, -1]]
            lo = np.argwhere(interp_points < x0)
            out[lo] = y0 + (y1 - y0) / (x1 - x0) * (interp_points[lo] - x0)
            hi = np.argwhere(interp_points > xn)
            out[hi] = yn + (yn - ym) / (xn - xm) * (interp_points[hi] - xn)
            return out
        if extrapolate == "constant":
            return np.interp(interp_points, fin_ref, fin_data, left=fin_data[0], right=fin_data[-1])
        raise ValueError(f"Unknown extrapolation method: {extrapolate}.")
    return np.interp(interp_points, fin_ref, fin_data, left=extrapolate[0], right=extrapolate[-1])


def eval_svd_model(svd_model, ass_ncoeff, param_list):
    """``nmma/em/lightcurve_generation.py:180-217``."""
    n_coeff = min(ass_ncoeff, svd_model["n_coeff"]) if ass_ncoeff else svd_model["n_coeff"]
    VA = svd_model["VA"]
    x = (np.array(param_list) - svd_model["param_mins"]) / (svd_model["param_maxs"] - svd_model["param_mins"])
    if "model" in svd_model:
        cAproj = svd_model["model"](np.atleast_2d(x)).numpy().T.flatten()
    else:
        cAproj = np.zeros((n_coeff,))
        gps = svd_model["gps"]
        if gps is None:
            raise ValueError("Gaussian process model unavailable.")
        for i in range(n_coeff):
            y_pred, _sigma = gps[i].predict(np.atleast_2d(x), return_std=True)
            cAproj[i] = y_pred.item()
    svd_back = np.dot(VA[:, :n_coeff], cAproj)
    svd_back *= svd_model["maxs"] - svd_model["mins"]
    svd_back += svd_model["mins"]
    return svd_model["tt"], svd_back


def calc_svd_lc(sample_times, param_list, svd_mag_model, mag_ncoeff=None, filters=None):
    """``nmma/em/lightcurve_generation.py:147-178``."""
    if filters is None:
        filters = list(svd_mag_model.keys())
    mAB = {f: np.full_like(sample_times, np.inf) for f in filters if f not in svd_mag_model}
    for filt in filters:
        if filt in mAB:
            continue
        tt, mag_back = eval_svd_model(svd_mag_model[filt], mag_ncoeff, param_list)
        mAB[filt] = autocomplete_data(sample_times, tt, mag_back, extrapolate=np.inf)
    return mAB


# --------------------------------------------------------------------------
# parameter conversion + detector-frame light curve
# --------------------------------------------------------------------------
def observation_angle_conversion(parameters):
    """``nmma/core/conversion.py:119-126``."""
    theta_jn = parameters.get("theta_jn", np.arccos(parameters.get("cos_theta_jn", 1.0)))
    theta_jn = np.minimum(theta_jn, np.pi - theta_jn)
    if "KNtheta" not in parameters:
        parameters["KNtheta"] = parameters.get("inclination_EM", theta_jn) * 180.0 / np.pi
    if "inclination_EM" not in parameters:
        parameters["inclination_EM"] = parameters["KNtheta"] / 180.0 * np.pi
    return parameters


def distance_modulus_nmma(d_lum=1e-5):
    """``nmma/core/conversion.py:30-34``."""
    return 5.0 * (5 + np.log10(d_lum))


class OracleSVDLightCurveModel:
    """``nmma/em/model.py:175-408`` (base container) + ``:535-731`` (SVD model)."""

    extinction_law = "P92_SMC_host"   # nmma/em/model.py:201

    def __init__(self, model_parameters, svd_mag_model, filters=None, sample_times=None,
                 mag_ncoeff=None, default_filts=None, lambdas=None, extinction_coef=None):
        # em/model.py:223-224: (default_filts, lambdas) = get_default_filts_lambdas(filters); nu_0s = c_SI / lambdas.
        # The table comes from sncosmo's bandpass registry in the reference; here it is handed in by the caller.
        self.default_filts = list(default_filts) if default_filts is not None else []
        self.nu_0s = 299792458.0 / np.asarray(lambdas, float) if lambdas is not None else np.zeros(0)
        self.extinction_coef = extinction_coef   # {filter: A_f / E(B-V)}: stands in for the G23 curve (dust_extinction absent)
        self.model_parameters = list(model_parameters)
        self.svd_mag_model = svd_mag_model
        self.filters = list(filters) if filters is not None else list(svd_mag_model.keys())
        self.mag_ncoeff = mag_ncoeff
        # em/model.py:230-232, 655-660
        self.model_times = (np.asarray(sample_times, float) if sample_times is not None
                            else next(iter(svd_mag_model.values()))["tt"])
        self.redshift_func = self._get_redshift
        self.good_parameters = True

    # nmma/core/conversion.py:57-64
    @staticmethod
    def _get_redshift(parameters):
        if "redshift" in parameters:
            return parameters["redshift"]
        if "luminosity_distance" in parameters:
            return _cosmo.Planck18.z_at_luminosity_distance(parameters["luminosity_distance"])
        return 0.0

    def check_vs_priors(self, dl_min=None, dl_max=None, has_redshift_prior=False, table=None):
        """``em/model.py:249-267``: dL -> z lookup on a 50-point grid."""
        if has_redshift_prior or (dl_min is None and table is None):
            return
        dist_grid, z_grid = table if table is not None else _cosmo.get_cosmo_grids(dl_min, dl_max)
        self.z_table = (np.asarray(dist_grid, float), np.asarray(z_grid, float))

        def redshift_from_dlum(parameters):
            return float(np.interp(parameters["luminosity_distance"], *self.z_table))

        self.redshift_func = redshift_from_dlum

    def parameter_conversion(self, parameters):
        """``em/model.py:272-286``."""
        new = observation_angle_conversion(parameters)
        for key in self.model_parameters:
            if key not in new:
                if key.lstrip("log10_") in new.keys():
                    new[key] = np.log10(new[key.lstrip("log10_")])
                elif "log10_" + key in new.keys():
                    new[key] = 10 ** new["log10_" + key]
        return new

    def em_parameter_setup(self, parameters):
        """``em/model.py:288-303``."""
        self.Ebv = parameters.get("Ebv", 0.0)
        self.luminosity_distance = parameters.get("luminosity_distance", 1e-5)
        self.distmod = distance_modulus_nmma(self.luminosity_distance)
        self.timeshift = parameters.get("timeshift", 0.0)
        self.redshift = self.redshift_func(parameters)
        return [parameters[k] if k in parameters else getattr(self, k) for k in self.model_parameters]

    def generate_lightcurve(self, sample_times, parameters, filters="all"):
        """``em/model.py:707-728``."""
        plist = self.em_parameter_setup(parameters)
        if filters == "all":
            filters = self.filters
        return calc_svd_lc(sample_times, plist, self.svd_mag_model,
                           mag_ncoeff=self.mag_ncoeff, filters=filters)

    def gen_detector_lc(self, parameters, sample_times=None):
        """``em/model.py:352-404`` with Ebv == 0 (extinction adds exactly 0.0)."""
        if sample_times is None:
            sample_times = self.model_times
        model_lc = self.generate_lightcurve(sample_times, parameters)
        observable_times = sample_times * (1 + self.redshift) + self.timeshift
        # get_extinction_mags + apply_extinction_correction, em/model.py:323-350,381-386
        coef = None
        if self.extinction_coef is not None:
            coef = [self.extinction_coef.get(f, 0.0) for f in self.default_filts]
        ext_mag = _ext.get_extinction_mags(self.nu_0s, self.Ebv, self.redshift, self.extinction_law, coef)
        for em, filt in zip(ext_mag, self.default_filts):
            if filt in model_lc:
                model_lc[filt] = model_lc[filt] + em
        redshift_correction = -2.5 * np.log10(1 + self.redshift)
        lc_data = {}
        for filt, mags in model_lc.items():
            if np.sum(np.isfinite(mags)) >= 2:
                lc_data[filt] = mags + self.distmod + redshift_correction
            else:
                lc_data[filt] = np.full_like(observable_times, np.inf)
        return observable_times, lc_data


# --------------------------------------------------------------------------
# filter-name mapping
# --------------------------------------------------------------------------
_HARDCODED = {"B": "g", "R": "z", "F160W": "H", "U": "u", "UVW2": "u", "UVW1": "u", "UVM2": "u"}
_AVERAGES = {"w": ["g", "r", "i"], "o": ["r", "i"], "c": ["g", "r"], "V": ["g", "r"],
             "F606W": ["g", "r"], "I": ["z", "y"], "F814W": ["z", "y"]}


def get_filter_name_mapping(observed_filters, known_filters):
    """``nmma/em/utils.py:478-563`` (``known_filters`` stands for the sncosmo registry)."""
    maps = {n: n for n in known_filters}
    maps.update(_HARDCODED)
    direct, averaging = {}, {}
    for f in observed_filters:
        if f in maps:
            direct[f] = maps[f]
        elif f.startswith("radio") or f.startswith("X-ray"):
            direct[f] = f
        elif f in _AVERAGES:
            averaging[f] = _AVERAGES[f]
        else:
            raise ValueError(f"Unknown filter: {f}. Cannot be processed")
    return direct, averaging


def average_mags(mag, filt):
    """``nmma/em/utils.py:566-584``."""
    if filt == "w":
        return (mag["g"] + mag["r"] + mag["i"]) / 3.0
    if filt in ["c", "V", "F606W"]:
        return (mag["g"] + mag["r"]) / 2.0
    if filt == "o":
        return (mag["r"] + mag["i"]) / 2.0
    if filt in ["I", "F814W"]:
        return (mag["z"] + mag["y"]) / 2.0
    raise ValueError(f"Unknown filter: {filt}")


# --------------------------------------------------------------------------
# systematics
# --------------------------------------------------------------------------
class OracleFilterSystematics:
    """Evaluation side of ``nmma/em/systematics.py:194-296``.

    ``budget``: {filt: float} constant error budget (``from_budget`` :51,203-210).
    ``direct``: {filt: param_name} (``from_param`` / ``from_single_params`` :279-286).
    ``interp``: {filt: (param_names, time_nodes)} (``from_interpolated_params`` :288-291).
    A filter listed in ``direct``/``interp`` overrides ``budget``; when any map is
    present only mapped filters are returned, as in the reference.
    """

    def __init__(self, filters, light_curve_times, budget=1.0, direct=None, interp=None):
        self.filters = list(filters)
        self.times = light_curve_times
        if not isinstance(budget, dict):
            budget = {f: float(budget) for f in self.filters}
        self.budget = budget
        self.direct = direct or {}
        self.interp = interp or {}

    def __call__(self, parameters):
        if not self.direct and not self.interp:
            return {f: np.full_like(self.times[f], self.budget[f]) for f in self.filters}
        out = {f: np.full_like(self.times[f], parameters[p]) for f, p in self.direct.items()}
        out.update({f: autocomplete_data(self.times[f], nodes, [parameters[p] for p in names],
                                         extrapolate="constant")
                    for f, (names, nodes) in self.interp.items()})
        return out


# --------------------------------------------------------------------------
# likelihood
# --------------------------------------------------------------------------
class OracleMultiFilterTransient:
    """``nmma/em/em_likelihood.py:136-352`` wrapped by ``nmma/core/base.py:77-82,178-182``."""

    def __init__(self, filters, light_curve_model, light_curve_data, systematics,
                 detection_limit=np.inf, known_filters=None, constraints=None):
        self.constraints = dict(constraints or {})   # {key: (minimum, maximum)} of the Constraint priors
        self.observed_filters = list(filters)
        known = set(known_filters) if known_filters is not None else set(light_curve_model.filters)
        self.model_filter_mapping, self.obs_average_mapping = get_filter_name_mapping(filters, known)
        self.light_curve_model = light_curve_model
        (self.light_curve_times, self.light_curves,
         self.light_curve_uncertainties, self.trigger_time) = light_curve_data
        self.systematics_handler = systematics
        if not isinstance(detection_limit, dict):
            detection_limit = {f: float(detection_limit) for f in self.observed_filters}
        self.detection_limit = {f: float(detection_limit.get(f, np.inf)) for f in self.observed_filters}

    # em_likelihood.py:305-311
    @staticmethod
    def sanity_check(model_lc):
        if not model_lc:
            return False
        if any([np.isinf(mag).all() for mag in model_lc.values()]):
            return False
        return True

    # em_likelihood.py:313-335
    def update_lightcurve_reference(self, obs_times, lc_data):
        expected = {}
        for filt in self.observed_filters:
            try:
                obs_mags = lc_data[self.model_filter_mapping[filt]]
                expected[filt] = autocomplete_data(self.light_curve_times[filt], obs_times, obs_mags,
                                                   extrapolate=np.inf)
            except KeyError:
                helper = {}
                for hf in self.obs_average_mapping[filt]:
                    obs_mags = lc_data[self.model_filter_mapping[hf]]
                    helper[hf] = autocomplete_data(self.light_curve_times[filt], obs_times, obs_mags,
                                                   extrapolate=np.inf)
                expected[filt] = average_mags(helper, filt)
        return expected

    # em_likelihood.py:224-256
    @staticmethod
    def chisquare_gaussianlog_from_lc_data(est_mag, data_mag, data_sigma, upperlim_sigma, lim=np.inf):
        finiteIdx = np.isfinite(data_sigma)
        infIdx = ~finiteIdx
        if finiteIdx.sum() >= 1:
            loc, scale = est_mag[finiteIdx], data_sigma[finiteIdx]
            b = (lim - loc) / scale
            minus_chisquare = np.sum(truncnorm.logpdf(data_mag[finiteIdx], -np.inf, b, loc=loc, scale=scale))
            if np.isnan(minus_chisquare):
                return False, -np.inf
        else:
            minus_chisquare = 0.0
        gausslogsf = np.zeros(2)
        if infIdx.sum() > 0:
            gausslogsf = norm.logsf(data_mag[infIdx], est_mag[infIdx], upperlim_sigma[infIdx])
        return minus_chisquare, np.sum(gausslogsf)

    # em_likelihood.py:337-352
    def band_log_likelihood(self, expected_mags, obs_error):
        chi_total, gauss_total = 0.0, 0.0
        for filt, err in obs_error.items():
            data_sigma = np.sqrt(self.light_curve_uncertainties[filt] ** 2 + err ** 2)
            chi, gauss = self.chisquare_gaussianlog_from_lc_data(
                expected_mags[filt], self.light_curves[filt], data_sigma, err,
                lim=self.detection_limit[filt])
            if chi is False:
                return SENTINEL
            chi_total += chi
            gauss_total += gauss
        return chi_total + gauss_total

    # em_likelihood.py:186-204
    def sub_log_likelihood(self, parameters):
        obs_times, model_lc = self.light_curve_model.gen_detector_lc(parameters)
        if not self.sanity_check(model_lc):
            return SENTINEL
        expected = self.update_lightcurve_reference(obs_times, model_lc)
        obs_error = self.systematics_handler(parameters)
        return self.band_log_likelihood(expected, obs_error)

    # core/base.py:67-68,77-82,178-182; bilby Constraint.prob = (val > minimum) & (val < maximum)
    def log_likelihood(self, parameters):
        with np.errstate(all="ignore"):
            parameters = self.light_curve_model.parameter_conversion(dict(parameters))
            ok = np.prod([(parameters[k] > lo) & (parameters[k] < hi) for k, (lo, hi) in self.constraints.items()])
            if not (ok and self.light_curve_model.good_parameters):
                return SENTINEL
            logl = self.sub_log_likelihood(parameters)
        if not np.isfinite(logl):
            return SENTINEL
        return float(logl)


def setup_filtered_lc_data(light_curve_data, trigger_time):
    """``nmma/em/utils.py:255-286``."""
    times, mags, errs = {}, {}, {}
    for filt, sub in light_curve_data.items():
        mags[filt] = np.array(sub["mag"])
        errs[filt] = np.array(sub["mag_error"])
        times[filt] = np.array(sub["time"]) - trigger_time
    return times, mags, errs, trigger_time
