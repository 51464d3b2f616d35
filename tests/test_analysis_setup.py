"""analysis_setup (nmma_b200/em/analysis.py, after nmma/em/analysis.py:110-173) from files in the reference's formats:
`time filter mag mag_error` rows with ISO times, a bilby-syntax .prior file, --data-tmax / --filters /
--remove-nondetections / --detection-limit.  No GPU needed: the engine is only built on the first evaluation."""
import json
import os

import numpy as np
import pytest

from nmma_b200 import synthetic as syn
from nmma_b200.em import analysis

PRIOR_TEXT = """luminosity_distance = Uniform(name='luminosity_distance', minimum=1, maximum=200)
KNphi = Uniform(name='KNphi', minimum=15., maximum=75.)
inclination_EM = Sine(name='inclination_EM', minimum=0., maximum=np.pi/2.)
timeshift = Uniform(name='timeshift', minimum=-2, maximum=0.1)
log10_mej_dyn = Uniform(name='log10_mej_dyn', minimum=-3., maximum=-1.)
log10_mej_wind = Uniform(name='log10_mej_wind', minimum=-3., maximum=-0.5)
"""


def _mjd_to_isot(mjd):
    import datetime
    t = datetime.datetime(1858, 11, 17) + datetime.timedelta(days=float(mjd))
    return t.strftime("%Y-%m-%dT%H:%M:%S.") + f"{t.microsecond // 1000:03d}"


@pytest.fixture()
def files(tmp_path):
    with open(os.path.join(os.path.dirname(syn.__file__), "data", "at2017gfo.json")) as fh:
        blob = json.load(fh)
    rows = []
    for filt, d in blob["data"].items():
        for t, m, e in zip(d["time"], d["mag"], d["mag_error"]):
            rows.append((float(t), f"{_mjd_to_isot(t)} {filt} {m} {'inf' if e == 'inf' else e}"))
    rows.sort()
    dat = tmp_path / "AT2017gfo.dat"
    dat.write_text("\n".join(r[1] for r in rows) + "\n")
    prior = tmp_path / "Bu2019lm.prior"
    prior.write_text(PRIOR_TEXT)
    return str(dat), str(prior), blob


def _args(dat, prior, *extra):
    return analysis.get_parser().parse_args(
        ["--model", "Bu2019lm", "--label", "x", "--light-curve-data", dat, "--prior", prior,
         "--trigger-time", str(syn.AT2017GFO_TRIGGER_MJD), "--em-error-budget", "1"] + list(extra))


def test_setup_from_reference_formats(files):
    dat, prior, blob = files
    filters = list(blob["data"])
    core = syn.random_model("Bu2019lm", filters, seed=0)
    priors, lik = analysis.analysis_setup(_args(dat, prior, "--data-tmax", "14"), svd_mag_model=core)
    # create_prior_from_args (nmma/em/prior.py:221-244) appends Ebv = DeltaFunction(0) when --use-Ebv is not given
    assert list(priors.keys()) == ["luminosity_distance", "KNphi", "inclination_EM", "timeshift", "log10_mej_dyn",
                                   "log10_mej_wind", "Ebv"]
    assert priors["Ebv"].peak == 0.0
    assert lik.columns == list(priors.keys())[:-1]
    _, lik_e = analysis.analysis_setup(_args(dat, prior, "--data-tmax", "14", "--use-Ebv", "--Ebv-max", "0.4"), svd_mag_model=core)
    assert lik_e.columns[-1] == "Ebv" and lik_e.priors["Ebv"].maximum == 0.4
    assert lik_e.priors["Ebv"].prob(0.0) == pytest.approx(2 / 0.4) and lik_e.priors["Ebv"].prob(0.4) == pytest.approx(0.0, abs=1e-12)
    plan = lik_e.sub_model.plan_layout(lik_e.columns)
    assert plan["ext"][0] == 1 and plan["ext"][1].col == 6 and np.all(plan["ext"][2] > 1e14)   # P92_SMC_host, nu0 in Hz
    sm = lik.sub_model
    times, mags, errs, trig = sm.light_curve_times, sm.light_curves, sm.light_curve_uncertainties, sm.trigger_time
    assert trig == syn.AT2017GFO_TRIGGER_MJD
    n = sum(len(v) for v in times.values())
    n_ul = sum(int(np.isinf(v).sum()) for v in errs.values())
    assert (n, n_ul) == (133, 3)                     # 133 rows <= 14 d, 3 upper limits
    ref, _ = syn.load_at2017gfo(data_tmax=14.0)      # the packaged copy through the same cuts
    for f in times:
        assert np.allclose(times[f], ref[0][f], atol=2e-8) and np.array_equal(mags[f], ref[1][f])   # ISO stamps carry ms


def test_setup_filters_nondetections_and_limits(files):
    dat, prior, blob = files
    core = syn.random_model("Bu2019lm", ["ps1::g", "ps1::i", "sdssu"], seed=0)
    _, lik = analysis.analysis_setup(
        _args(dat, prior, "--data-tmax", "10", "--filters", "ps1::g,ps1::i,sdssu", "--remove-nondetections",
              "--detection-limit", "24.5"), svd_mag_model=core)
    sm = lik.sub_model
    times, errs = sm.light_curve_times, sm.light_curve_uncertainties
    assert set(times) <= {"ps1::g", "ps1::i", "sdssu"}
    assert all(np.all(np.isfinite(e)) for e in errs.values())          # upper limits removed
    assert all(np.all(t <= 10.0) and np.all(t >= 0.0) for t in times.values())
    assert all(lik.sub_model.detection_limit[f] == 24.5 for f in times)


def test_setup_rejects_data_outside_the_model_window(files):
    dat, prior, blob = files
    core = syn.random_model("Bu2019lm", list(blob["data"]), seed=0)
    with pytest.raises(ValueError, match="Last data point"):
        analysis.analysis_setup(_args(dat, prior), svd_mag_model=core)   # data reach 25.4 d, the model grid ends at 21 d
