"""Observation readers (``nmma/em/io.py:16-144``) without astropy.

File format: whitespace-separated ``time filter mag mag_error`` rows; ``time`` is an ISO-8601
UTC string or a float MJD (``--time-format mjd``, the reference default) ; ``mag_error = inf``
marks an upper limit.  JSON files in the standard ``{filter: {time, mag, mag_error}}`` layout
are accepted as well.
"""
from __future__ import annotations

import argparse
import json
from datetime import datetime, timezone

import numpy as np

_MJD_EPOCH = datetime(1858, 11, 17, tzinfo=timezone.utc)


def isot_to_mjd(stamp: str) -> float:
    """UTC ISO-8601 -> MJD (``astropy.time.Time(stamp).mjd`` for the utc scale; leap-second days
    are stretched by astropy, a < 1.2e-5 d effect that no shipped light curve hits)."""
    s = stamp.strip().replace("Z", "")
    if "T" not in s and " " in s:
        s = s.replace(" ", "T")
    if "." in s:
        head, frac = s.split(".")
        frac = (frac + "000000")[:6]
        s = f"{head}.{frac}"
    dt = datetime.fromisoformat(s)
    if dt.tzinfo is None:
        dt = dt.replace(tzinfo=timezone.utc)
    delta = dt - _MJD_EPOCH
    return delta.days + (delta.seconds + delta.microseconds * 1e-6) / 86400.0


def parse_time(token: str, time_format=None) -> float:
    """``Time(token).mjd`` with the reference's fallback to ``Time(token, format=time_format or 'mjd')``."""
    try:
        return isot_to_mjd(token)
    except ValueError:
        fmt = time_format or "mjd"
        val = float(token)
        if fmt == "mjd":
            return val
        if fmt == "jd":
            return val - 2400000.5
        if fmt == "gps":
            return gps_to_mjd(val)
        raise ValueError(f"unsupported time format {fmt!r}")


# TAI-UTC steps after the GPS epoch (1980-01-06, TAI-UTC = 19 s): MJD of the day each leap second took effect
_LEAP_MJDS = [44786, 45151, 45516, 46247, 47161, 47892, 48257, 48804, 49169, 49534, 50083, 50630, 51179,
              53736, 54832, 56109, 57204, 57754]


def gps_to_mjd(gps_seconds: float) -> float:
    """GPS seconds -> UTC MJD (``Time(gps, format='gps').mjd``)."""
    mjd = 44244.0 + gps_seconds / 86400.0
    for _ in range(2):
        nleap = sum(1 for m in _LEAP_MJDS if mjd >= m)
        mjd = 44244.0 + (gps_seconds - nleap) / 86400.0
    return mjd


def strict_read_csv(filename, args=None):
    """``nmma/em/io.py:116-144``."""
    data = {}
    time_format = getattr(args, "time_format", None)
    with open(filename, "r") as fh:
        for line in fh:
            line = line.rstrip("\n")
            if not line or line.startswith("#") or line.startswith("time") or line.startswith("mjd"):
                continue
            parts = line.split(None)
            mjd = parse_time(parts[0], time_format)
            filt, mag, dmag = parts[1], float(parts[2]), float(parts[3])
            entry = data.setdefault(filt, {"time": [], "mag": [], "mag_error": []})
            entry["time"].append(mjd)
            entry["mag"].append(mag)
            entry["mag_error"].append(dmag)
    return data


def read_lc_from_json(filename):
    """``nmma/em/io.py:62-80`` (standard or 'model' layout)."""
    with open(filename, "r") as fh:
        data = json.load(fh)
    if "time" in data:
        new = {}
        for key, value in data.items():
            if key != "time" and not key.endswith("_error"):
                new[key] = {"time": data["time"], "mag": value,
                            "mag_error": data.get(f"{key}_error", np.zeros_like(data["time"]))}
        data = new
    return data


def load_em_observations(filename, args=None, format="observations"):
    """``nmma/em/io.py:16-59``: -> {filter: {'time','mag','mag_error'} arrays} (times in MJD)."""
    if isinstance(filename, dict):
        return filename
    if isinstance(filename, argparse.Namespace):
        args = filename
        filename = args.light_curve_data
    if isinstance(filename, dict):
        return filename
    if filename is None:
        raise ValueError("No filename provided for lightcurve data.")
    if filename.endswith(".json"):
        data = read_lc_from_json(filename)
    elif "obs" in format:
        data = strict_read_csv(filename, args)
    else:
        raise ValueError(f"format {format!r} is not supported by nmma_b200 (observations / json only)")
    return {filt: {k: np.array(v) for k, v in d.items()} for filt, d in data.items()}
