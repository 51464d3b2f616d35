"""Host-side setup helpers of the hot path (``nmma/em/utils.py:72-349, 478-593, 626-677``).

These run once per analysis (filters, detection limits, data dict -> per-filter
arrays, model/data time-range check, filter-name mapping).  Per-point arithmetic
(``autocomplete_data`` on the light curve) lives in the CUDA back end; the NumPy
``autocomplete_data`` here serves one-off host uses (systematics preview, data prep).
"""
from __future__ import annotations

import numpy as np

from ..core.conversion import luminosity_distance_to_redshift

# sncosmo's registered bandpass names that NMMA models use.  The reference builds this
# list from sncosmo's registry (``nmma/em/utils.py:470-476``); sncosmo is absent here, so
# the names are listed (default ``filts`` of em/utils.py:40-69 + ALLOWED_FILTERS of
# em/systematics.py:343-370 + the LSST/ZTF/UVOT/HST families).
SNCOSMO_BANDPASSES = [
    "bessellux", "bessellb", "bessellv", "bessellr", "besselli",
    "standard::u", "standard::b", "standard::v", "standard::r", "standard::i",
    "desu", "desg", "desr", "desi", "desz", "desy",
    "sdssu", "sdssg", "sdssr", "sdssi", "sdssz",
    "sdss::u", "sdss::g", "sdss::r", "sdss::i", "sdss::z",
    "ps1::g", "ps1::r", "ps1::i", "ps1::z", "ps1::y", "ps1::open", "ps1::w",
    "lsstu", "lsstg", "lsstr", "lssti", "lsstz", "lssty",
    "ztfg", "ztfr", "ztfi", "ztf::g", "ztf::r", "ztf::i",
    "2massj", "2massh", "2massks", "atlasc", "atlaso",
    "uvot::b", "uvot::u", "uvot::uvm2", "uvot::uvw1", "uvot::uvw2", "uvot::v", "uvot::white",
    "cspb", "csphs", "csphd", "cspjs", "cspjd", "cspv3009", "cspv3014", "cspv9844",
    "cspys", "cspyd", "cspg", "cspi", "cspk", "cspr", "cspu",
    "f435w", "f475w", "f555w", "f606w", "f625w", "f775w", "f814w", "f850lp",
    "f105w", "f110w", "f125w", "f127m", "f139m", "f140w", "f153m", "f160w",
    "f070w", "f090w", "f115w", "f150w", "f200w", "f277w", "f356w", "f444w",
    "f560w", "f770w", "f1000w", "f1130w", "f1280w", "f1500w", "f1800w", "f2100w", "f2550w",
    "gaia::g", "gaia::gbp", "gaia::grp", "gaia::grvs", "kepler", "tess", "ultrasat",
    "gotob", "gotog", "gotol", "gotor", "galex::fuv", "galex::nuv",
    "swope2::u", "swope2::b", "swope2::g", "swope2::v", "swope2::v1", "swope2::v2", "swope2::r",
    "swope2::i", "swope2::y", "swope2::J", "swope2::H",
]

_UNPROCESSED = ["u", "g", "r", "i", "z", "y", "J", "H", "K", "X-ray-1keV", "X-ray-5keV",
                "radio-5.5GHz", "radio-1.25GHz", "radio-6GHz", "radio-3GHz",
                "sdss::u", "sdss::g", "sdss::r", "sdss::i", "sdss::z", "swope2::y", "swope2::J", "swope2::H"]
_HARDCODED = {"B": "g", "R": "z", "F160W": "H", "U": "u", "UVW2": "u", "UVW1": "u", "UVM2": "u"}


_WAVE_EFF = None


def _wave_eff_table():
    """{sncosmo bandpass name: wave_eff [Angstrom]} from ``nmma_b200/data/wave_eff.json`` (tools/make_wave_eff.py:
    sncosmo's definition evaluated on the transmission tables vendored in nmma-data; reproduces the six PS1 values the
    reference hard-codes in ``lambdas_sloan``, nmma/em/utils.py:712-714, to the printed digit)."""
    global _WAVE_EFF
    if _WAVE_EFF is None:
        import json
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "wave_eff.json")
        with open(path) as fh:
            _WAVE_EFF = json.load(fh)["wave_eff"]
    return _WAVE_EFF


def get_default_filts_lambdas(filters=None):
    """``nmma/em/utils.py:680-779``: (filter names, effective wavelengths in metres).  Filters that are neither in the
    hard-coded list nor a known bandpass are dropped with the reference's warning (they stay uncorrected for extinction)."""
    from ..core.constants import c_SI
    eV_per_h_SI = 2.417989242e14          # e / h [Hz per eV] (nmma/core/constants.py)
    filts = ["u", "g", "r", "i", "z", "y", "J", "H", "K", "U", "B", "V", "R", "I",
             "radio-1.25GHz", "radio-3GHz", "radio-5.5GHz", "radio-6GHz", "X-ray-1keV", "X-ray-5keV"]
    lambdas = list(1e-10 * np.array([3561.8, 4866.46, 6214.6, 7687.0, 7127.0, 7544.6, 8679.5, 9633.3, 12350.0]))
    lambdas += list(1e-10 * np.array([3605.07, 4413.08, 5512.12, 6585.91, 8059.88]))
    lambdas += list(c_SI / np.array([1.25e9, 3e9, 5.5e9, 6e9]))
    lambdas += list(c_SI / (np.array([1e3, 5e3]) * eV_per_h_SI))
    # the reference's list also names sdss::*, swope2::*, FUV, NUV without a wavelength entry (its `filts` is longer than
    # `lambdas` there, which shifts nothing because lookups go through the bandpass entries appended below)
    for name, w in _wave_eff_table().items():
        filts.append(name)
        lambdas.append(1e-10 * w)
    if filters is None:
        return filts, np.array(lambdas)
    out_f, out_l = [], []
    for filt in filters:
        if filt.startswith("radio") and filt not in filts:
            unit = {"GHz": 1e9, "MHz": 1e6, "kHz": 1e3}.get(filt[-3:])
            out_f.append(filt); out_l.append(c_SI / (float(filt.replace("radio-", "")[:-3]) * unit))
        elif filt.startswith("X-ray-") and filt not in filts:
            unit = {"keV": 1e3, "MeV": 1e6}.get(filt[-3:])
            out_f.append(filt); out_l.append(c_SI / (float(filt.replace("X-ray-", "")[:-3]) * unit * eV_per_h_SI))
        elif filt in filts:
            ii = filts.index(filt)
            out_f.append(filts[ii]); out_l.append(lambdas[ii])
        else:
            print(f"Warning: {filt} not found in filter list.")
    return out_f, np.array(out_l)


def setup_sample_times(args):
    """``nmma/em/utils.py:72-93``."""
    tmin, tmax = args.em_tmin, args.em_tmax
    if tmin is None and tmax is None:
        return None
    if getattr(args, "em_tstep", None):
        return np.arange(tmin, tmax + args.em_tstep, args.em_tstep)
    timescale = getattr(args, "em_timescale", "linear")
    nsteps = getattr(args, "em_nsteps", 150)
    if "lin" in timescale or tmin <= 0.0:
        return np.linspace(tmin, tmax, nsteps)
    if any(s in timescale for s in ["log", "geo"]):
        return np.geomspace(tmin, tmax, nsteps)
    raise ValueError(f"Unknown time scale {timescale}. Please use 'lin(ear)' or 'log(arithmic)' / 'geo(metric)'.")


def set_filters(args):
    """``nmma/em/utils.py:96-139`` (explicit ``--filters`` and the ztf/rubin/lsst detector shortcuts)."""
    filters = None
    if getattr(args, "filters", None):
        filters = args.filters
        if isinstance(filters, str):
            filters = filters.split(",")
        filters = [f.replace(" ", "").split(",") for f in filters]
        filters = [f for sub in filters for f in sub if f]
        if len(filters) == 0:
            raise ValueError("Need at least one valid filter.")
    elif getattr(args, "em_detectors", None) or getattr(args, "rubin_ToO_type", False):
        dets = args.em_detectors.split(",") if isinstance(args.em_detectors, str) else list(getattr(args, "em_detectors", None) or [])
        dets = [d.strip().lower() for d in dets]
        filters = []
        if "ztf" in dets:
            dets.remove("ztf")
            filters.extend(["ztfg", "ztfr", "ztfi"])
        if "lsst" in dets:
            dets.remove("lsst")
            filters.extend(["lsstg", "lsstr", "lssti", "lsstz", "lssty"])
        elif getattr(args, "rubin_ToO_type", None):
            table = {"platinum": "grizy", "gold": "gri", "gold_z": "grz", "silver": "gi", "silver_z": "gz"}
            filters.extend([f"ps1::{b}" for b in table.get(args.rubin_ToO_type, "")])
            if "rubin" in dets:
                dets.remove("rubin")
        elif "rubin" in dets:
            dets.remove("rubin")
            filters.extend(["ps1::g", "ps1::r", "ps1::i", "ps1::z", "ps1::y"])
        if dets:
            raise NotImplementedError(f"{dets} not implemented yet.")
    return filters


def set_filter_associated_dict(quantity, filters, default_limit=np.inf):
    """``nmma/em/utils.py:213-230``."""
    if isinstance(quantity, (int, float)):
        return {x: float(quantity) for x in filters}
    if isinstance(quantity, (list, tuple)):
        assert len(quantity) == len(filters), f" {quantity} must match the number of filters: {filters}."
        return {x: float(y) for x, y in zip(filters, quantity)}
    if isinstance(quantity, dict):
        return {filt: float(quantity.get(filt, default_limit)) for filt in filters}
    raise ValueError(f"Could not derive a dict for {quantity} and filters {filters}.")


def create_detection_limit(args, filters, default_limit=np.inf):
    """``nmma/em/utils.py:142-195`` (FITS-map variant out of scope)."""
    if getattr(args, "detection_limit", None):
        return set_filter_associated_dict(args.detection_limit, filters, default_limit)
    if getattr(args, "detection_limit_fits_file", None):
        raise NotImplementedError("--detection-limit-fits-file (healpy/astropy) is outside the nmma_b200 hot path")
    detection_limit = {filt: default_limit for filt in filters}
    dets = getattr(args, "em_detectors", None)
    if dets:
        dets = dets.split(",") if isinstance(dets, str) else list(dets)
        if "lsst" in dets:
            dets.remove("lsst")
            detection_limit.update({"lsstu": 23.9, "lsstg": 25.0, "lsstr": 24.7, "lssti": 24.0, "lsstz": 23.3, "lssty": 22.1})
        if "ztf" in dets:
            dets.remove("ztf")
            detection_limit.update({"ztfg": 21.7, "ztfr": 21.4, "ztfi": 20.9})
        if "rubin" in dets:
            dets.remove("rubin")
            detection_limit.update({"ps1::g": 25.8, "ps1::r": 25.5, "ps1::i": 24.8, "ps1::z": 24.1, "ps1::y": 22.9})
        if dets:
            raise NotImplementedError(f"{dets} not implemented yet.")
    if getattr(args, "rubin_ToO_type", None):
        detection_limit.update({"ps1::g": 25.8, "ps1::r": 25.5, "ps1::i": 24.8, "ps1::z": 24.1, "ps1::y": 22.9})
    return detection_limit


def cut_data_to_time_range(data, args, trigger_time, tmin=0, tmax=np.inf):
    """``nmma/em/utils.py:233-252`` (``--data-tmin/--data-tmax``)."""
    tmin = getattr(args, "data_tmin", tmin)
    tmax = getattr(args, "data_tmax", tmax)
    tmin = 0 if tmin is None else tmin
    tmax = np.inf if tmax is None else tmax
    for filt in list(data.keys()):
        detector_time = data[filt]["time"] - trigger_time
        mask = (tmin <= detector_time) & (detector_time <= tmax)
        if not np.any(mask):
            del data[filt]
        else:
            data[filt] = {k: data[filt][k][mask] for k in ("time", "mag", "mag_error")}
    return data


def setup_filtered_lc_data(light_curve_data, trigger_time):
    """``nmma/em/utils.py:255-286``: (times - trigger, mags, errs, trigger) per filter."""
    lc_times, lc_mags, lc_unc = {}, {}, {}
    min_time = np.inf
    for filt, sub in light_curve_data.items():
        lc_mags[filt] = np.array(sub["mag"])
        lc_unc[filt] = np.array(sub["mag_error"])
        lc_times[filt] = np.array(sub["time"])
        min_time = np.minimum(min_time, np.min(sub["time"]))
    if min_time < 0:
        raise ValueError(f"trigger_time is {-min_time} days later than earliest data time. "
                         "Please provide a valid trigger time.")
    lc_times = {filt: lc_times[filt] - trigger_time for filt in lc_times}
    return (lc_times, lc_mags, lc_unc, trigger_time)


def check_model_time_consistency(light_curve_data, light_curve_model, priors, injection=None):
    """``nmma/em/utils.py:289-349``: raise if a detection can fall outside the model's
    guaranteed detector-frame window given the prior extremes of redshift and timeshift."""
    lc_times, lc_mags, lc_unc, trigger_time = light_curve_data
    data_tmin, data_tmax = np.inf, -np.inf
    for key in lc_times:
        det = np.isfinite(lc_mags[key]) & np.isfinite(lc_unc[key])
        data_tmin = np.minimum(data_tmin, lc_times[key][det].min())
        data_tmax = np.maximum(data_tmax, lc_times[key][det].max())
    zmin = zmax = 0.0
    if "redshift" in priors:
        zmin, zmax = priors["redshift"].minimum, priors["redshift"].maximum
    elif "luminosity_distance" in priors:
        if "Hubble_constant" in priors:
            raise NotImplementedError("Hubble_constant sampling is outside the nmma_b200 hot path")
        zmin = luminosity_distance_to_redshift(priors["luminosity_distance"].minimum)
        zmax = luminosity_distance_to_redshift(priors["luminosity_distance"].maximum)
    try:
        t0_min, t0_max = priors["timeshift"].minimum, priors["timeshift"].maximum
    except KeyError:
        t0_min, t0_max = 0.0, 0.0
    t_source_min, t_source_max = np.asarray(light_curve_model.model_times)[[0, -1]]
    t_obs_start_max = (1 + zmax) * t_source_min + t0_max
    t_obs_end_min = (1 + zmin) * t_source_max + t0_min
    if injection is not None:
        for key, time in lc_times.items():
            use = (time >= t_obs_start_max) & (time <= t_obs_end_min)
            lc_times[key], lc_mags[key], lc_unc[key] = time[use], lc_mags[key][use], lc_unc[key][use]
    elif data_tmin < t_obs_start_max:
        raise ValueError(f"First data point is at {data_tmin} days, but with your timeshift and redshift settings, "
                         f"the model time in detector frame can start as late as {t_obs_start_max}.")
    elif t_obs_end_min < data_tmax:
        raise ValueError(f"Last data point is at {data_tmax} days, but with your timeshift and redshift settings, "
                         f"the model time in detector frame can end as early as {t_obs_end_min}.")
    return (lc_times, lc_mags, lc_unc, trigger_time)


def map_observable_to_modelled_filters(obs_filter):
    """``nmma/em/utils.py:549-563``."""
    map_dict = {"w": ["g", "r", "i"], "o": ["r", "i"]}
    for f in ["c", "V", "F606W"]:
        map_dict[f] = ["g", "r"]
    for f in ["I", "F814W"]:
        map_dict[f] = ["z", "y"]
    if obs_filter in map_dict:
        return map_dict[obs_filter]
    raise ValueError(f"Unknown filter: {obs_filter}. Cannot be processed")


def get_filter_name_mapping(observed_filters, extra_known=()):
    """``nmma/em/utils.py:478-546``.  ``extra_known``: names the light-curve model itself provides,
    accepted as direct maps even when they are not in the static sncosmo name list."""
    maps = {n: n for n in _UNPROCESSED + SNCOSMO_BANDPASSES + list(extra_known)}
    maps.update(_HARDCODED)
    direct, averaging = {}, {}
    if isinstance(observed_filters, str):
        observed_filters = [observed_filters]
    for f in observed_filters:
        if f in maps:
            direct[f] = maps[f]
        elif f.startswith("radio") or f.startswith("X-ray"):
            direct[f] = f
        else:
            averaging[f] = map_observable_to_modelled_filters(f)
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


def autocomplete_data(interp_points, ref_points, ref_data, extrapolate="linear", ref_value=np.inf):
    """``nmma/em/utils.py:626-677`` for one-off host use (not on the per-point path)."""
    mask = np.isfinite(ref_data)
    if np.sum(mask) < 2:
        return np.full_like(interp_points, ref_value)
    xr = np.asarray(ref_points)[mask]
    yr = np.asarray(ref_data)[mask]
    x = np.atleast_1d(interp_points)
    if isinstance(extrapolate, (float, int)):
        return np.interp(x, xr, yr, left=extrapolate, right=extrapolate)
    if isinstance(extrapolate, str):
        if extrapolate == "linear":
            out = np.interp(x, xr, yr)
            lo, hi = x < xr[0], x > xr[-1]
            out[lo] = yr[0] + (yr[1] - yr[0]) / (xr[1] - xr[0]) * (x[lo] - xr[0])
            out[hi] = yr[-1] + (yr[-1] - yr[-2]) / (xr[-1] - xr[-2]) * (x[hi] - xr[-1])
            return out
        if extrapolate == "constant":
            return np.interp(x, xr, yr, left=yr[0], right=yr[-1])
        raise ValueError(f"Unknown extrapolation method: {extrapolate}.")
    return np.interp(x, xr, yr, left=extrapolate[0], right=extrapolate[-1])
