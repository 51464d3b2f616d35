#!/usr/bin/env python
"""Effective wavelengths of the photometric filters, for the extinction laws (nmma/em/utils.py:680-779).

The reference takes ``1e-10 * sncosmo.get_bandpass(name).wave_eff`` for every registered bandpass
(``get_default_filts_lambdas``, nmma/em/utils.py:725-736).  sncosmo is not installed offline, but its transmission tables
are vendored under ``/root/reference/nmma-data/sncosmo/bandpasses``; this script restates sncosmo's definition

    wave_eff = sum(w T(w)) / sum(T(w))   on the mid-point grid of spacing <= 5 A between the first and last table entry,

with T the piecewise-linear interpolant of the table after trimming leading / trailing entries below 1e-3 of the peak
(one entry kept on each side), which is how sncosmo loads its built-in tables -- restated from memory of
sncosmo/bandpasses.py (Bandpass.__init__, Bandpass.wave_eff, integration_grid), NOT checked against sncosmo:
"parity unpinned" (DESIGN.md).  Output: nmma_b200/data/wave_eff.json {filter name: wave_eff in Angstrom}.

    python tools/make_wave_eff.py [/root/reference/nmma-data/sncosmo/bandpasses]
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/nmma-data/sncosmo/bandpasses"

# sncosmo name -> (file relative to SRC, factor that converts the table's wavelength column to Angstrom)
TABLES = {
    "ps1::g": ("ps1/ps1_g.dat", 10.0), "ps1::r": ("ps1/ps1_r.dat", 10.0), "ps1::i": ("ps1/ps1_i.dat", 10.0),
    "ps1::z": ("ps1/ps1_z.dat", 10.0), "ps1::y": ("ps1/ps1_y.dat", 10.0), "ps1::w": ("ps1/ps1_w.dat", 10.0),
    "ps1::open": ("ps1/ps1_open.dat", 10.0),
    "2massj": ("2mass/2mass.J", 1.0), "2massh": ("2mass/2mass.H", 1.0), "2massks": ("2mass/2mass.Ks", 1.0),
    "sdssu": ("sdss/sdss_u.dat", 1.0), "sdssg": ("sdss/sdss_g.dat", 1.0), "sdssr": ("sdss/sdss_r.dat", 1.0),
    "sdssi": ("sdss/sdss_i.dat", 1.0), "sdssz": ("sdss/sdss_z.dat", 1.0),
    "ztfg": ("ztf/P48_g.dat", 1.0), "ztfr": ("ztf/P48_R.dat", 1.0), "ztfi": ("ztf/P48_I.dat", 1.0),
    "atlasc": ("atlas/Atlas.Cyan", 1.0), "atlaso": ("atlas/Atlas.Orange", 1.0),
    "uvot::b": ("swift/Swift_UVOT.B.dat", 1.0), "uvot::u": ("swift/Swift_UVOT.U.dat", 1.0),
    "uvot::uvm2": ("swift/Swift_UVOT.UVM2.dat", 1.0), "uvot::uvw1": ("swift/Swift_UVOT.UVW1.dat", 1.0),
    "uvot::uvw2": ("swift/Swift_UVOT.UVW2.dat", 1.0), "uvot::v": ("swift/Swift_UVOT.V.dat", 1.0),
    "uvot::white": ("swift/Swift_UVOT.white.dat", 1.0),
    "lsstu": ("lsst/total_u.dat", 10.0), "lsstg": ("lsst/total_g.dat", 10.0), "lsstr": ("lsst/total_r.dat", 10.0),
    "lssti": ("lsst/total_i.dat", 10.0), "lsstz": ("lsst/total_z.dat", 10.0), "lssty": ("lsst/total_y.dat", 10.0),
}
SPACING = 5.0       # sncosmo.constants.MODEL_BANDFLUX_SPACING
TRIM_LEVEL = 1e-3   # sncosmo built-in loaders


def read_table(path, factor):
    rows = []
    with open(path) as fh:
        for line in fh:
            line = line.split("#")[0].strip()
            if not line:
                continue
            parts = line.replace(",", " ").split()
            try:
                rows.append((float(parts[0]) * factor, float(parts[1])))
            except (ValueError, IndexError):
                continue
    a = np.array(rows, float)
    return a[:, 0], a[:, 1]


def trim(wave, trans, level):
    """slice_exclude_below(trans, max * level, grow=1)."""
    keep = np.nonzero(trans >= trans.max() * level)[0]
    i0, i1 = max(keep[0] - 1, 0), min(keep[-1] + 2, len(trans))
    return wave[i0:i1], trans[i0:i1]


def wave_eff(wave, trans):
    lo, hi = wave[0], wave[-1]
    nbin = int(math.ceil((hi - lo) / SPACING))
    step = (hi - lo) / nbin
    grid = np.linspace(lo + 0.5 * step, hi - 0.5 * step, nbin)
    w = np.interp(grid, wave, trans)
    return float(np.sum(grid * w) / np.sum(w))


def main():
    out = {}
    for name, (rel, factor) in sorted(TABLES.items()):
        path = os.path.join(SRC, rel)
        if not os.path.isfile(path):
            print(f"skip {name}: {path} not found")
            continue
        wave, trans = read_table(path, factor)
        if wave[0] < 200.0 or wave[-1] > 1e6:
            raise SystemExit(f"{name}: wavelengths {wave[0]}..{wave[-1]} A look like the wrong unit")
        wave, trans = trim(wave, trans, TRIM_LEVEL)
        out[name] = round(wave_eff(wave, trans), 4)
        print(f"{name:14s} {out[name]:10.2f} A  ({len(wave)} nodes)")
    dst = os.path.join(ROOT, "nmma_b200", "data", "wave_eff.json")
    with open(dst, "w") as fh:
        json.dump({"source": "tools/make_wave_eff.py from nmma-data/sncosmo/bandpasses (sncosmo wave_eff definition restated; "
                             "not checked against sncosmo)", "unit": "Angstrom", "wave_eff": out}, fh, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
