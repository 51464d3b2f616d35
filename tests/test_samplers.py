"""Batched nested sampler (nmma_b200/samplers.py) on likelihoods with known evidence, and the lightcurve-analysis
driver end to end on the GPU likelihood."""
import json
import os

import numpy as np
import pytest

from nmma_b200.samplers import equal_weight, nested_sample


@pytest.mark.parametrize("ndim,sigma", [(3, 0.08), (6, 0.03)])
def test_gaussian_evidence_and_posterior(ndim, sigma):
    mu = np.linspace(0.35, 0.65, ndim)

    def loglike(u):
        return -0.5 * np.sum(((u - mu) / sigma) ** 2, axis=1)

    truth = ndim * np.log(sigma * np.sqrt(2 * np.pi))        # unit-cube prior, Gaussian far from the walls
    res = nested_sample(loglike, ndim, nlive=400, batch=2048, dlogz=0.05, seed=3)
    assert abs(res["log_evidence"] - truth) < 3.5 * res["log_evidence_err"] + 0.05
    post = res["samples_u"][equal_weight(res, 4000, seed=1)]
    assert np.abs(post.mean(axis=0) - mu).max() < 4 * sigma / np.sqrt(400)
    assert np.abs(post.std(axis=0) / sigma - 1).max() < 0.15
    assert np.isclose(np.exp(res["log_weights"]).sum(), 1.0)


def test_sentinel_rows_count_as_zero_likelihood():
    # half of the prior returns the reference's failure sentinel: the evidence halves, nothing breaks
    def loglike(u):
        out = -0.5 * np.sum(((u - 0.5) / 0.1) ** 2, axis=1)
        out[u[:, 0] > 0.5] = -1.7976931348623157e308
        return out

    res = nested_sample(loglike, 2, nlive=300, batch=1024, dlogz=0.05, seed=5)
    truth = np.log(0.5 * 2 * np.pi * 0.1 ** 2)
    assert abs(res["log_evidence"] - truth) < 3.5 * res["log_evidence_err"] + 0.1
    assert np.all(res["samples_u"][np.isfinite(res["log_likelihoods"]), 0] <= 0.5)


def test_driver_parser_takes_the_reference_flags():
    """Option names and dests of nmma/em/em_parsing.py that reach the kilonova likelihood."""
    from nmma_b200.em import analysis
    a = analysis.get_parser().parse_args(
        ["--model", "Bu2019lm", "--svd-path", "svdmodels", "--interpolation-type", "tensorflow", "--outdir", "o",
         "--label", "l", "--trigger-time", "57982.5285236896", "--data", "AT2017gfo.dat", "--prior", "Bu2019lm.prior",
         "--tmin", "0.1", "--tmax", "14", "--dt", "0.5", "--error-budget", "1", "--nlive", "256", "--filters", "ps1::g,ps1::r",
         "--detection-limit", "24.5", "--remove-nondetections", "--svd-mag-ncoeff", "10", "--data-tmax", "14"])
    assert a.em_model == "Bu2019lm" and a.light_curve_data == "AT2017gfo.dat" and a.em_tmin == 0.1 and a.em_tmax == 14
    assert a.em_tstep == 0.5 and a.em_error_budget == 1.0 and a.nlive == 256 and a.detection_limit == 24.5
    b = analysis.get_parser().parse_args(["--kilonova-model", "Ka2017", "--label", "l", "--light-curve-data", "x",
                                          "--prior", "p", "--kilonova-tmin", "0.2", "--em-tstep", "0.1", "--kilonova-error", "0.5"])
    assert b.em_model == "Ka2017" and b.em_tmin == 0.2 and b.em_tstep == 0.1 and b.em_error_budget == 0.5


@pytest.mark.gpu
def test_lightcurve_analysis_driver(tmp_path):
    """Config 1 end to end on the device: AT2017gfo vs a Bu2019lm-shaped surrogate, nested sampling through the batched
    likelihood; ln Z is checked against a brute-force prior sweep (1e8 Philox draws scored on the device)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from nmma_b200 import synthetic as syn
    from nmma_b200.em import analysis

    with open(os.path.join(os.path.dirname(syn.__file__), "data", "at2017gfo.json")) as fh:
        blob = json.load(fh)
    data = {f: {k: np.array([np.inf if x == "inf" else x for x in v], float) for k, v in d.items()}
            for f, d in blob["data"].items()}
    # a 3 mag error budget keeps the posterior broad enough for the brute-force sweep to have thousands of effective
    # samples (a random-init surrogate does not fit AT2017gfo; with 1 mag the sweep's evidence is the noisier number)
    args = analysis.get_parser().parse_args(
        ["--model", "Bu2019lm", "--label", "at2017gfo", "--outdir", str(tmp_path), "--light-curve-data", "unused",
         "--prior", "unused", "--trigger-time", str(syn.AT2017GFO_TRIGGER_MJD), "--data-tmax", "14", "--nlive", "500",
         "--em-error-budget", "3", "--batch", "65536", "--max-calls", "30000000", "--dlogz", "0.2"])
    args.light_curve_data = data                      # load_em_observations accepts the dict form
    args.prior = syn.bu2019lm_prior()
    core = syn.random_model("Bu2019lm", list(data), seed=0)
    res = analysis.analysis(args, svd_mag_model=core)
    assert np.isfinite(res["log_evidence"]) and res["log_evidence_err"] < 0.3
    assert os.path.isfile(tmp_path / "at2017gfo_result.json") and os.path.isfile(tmp_path / "at2017gfo_posterior_samples.dat")
    assert set(res["search_parameter_keys"]) <= set(res["bestfit_params"])

    _, lik = analysis.analysis_setup(args, svd_mag_model=core)
    tot, n, best = -np.inf, 0, -np.inf
    for blk in range(10):
        out = lik.log_likelihood_sweep(10_000_000, seed=7, first_index=blk * 10_000_000)
        o = out.cpu().numpy() if hasattr(out, "cpu") else np.asarray(out)
        o = o[o > -1e300]
        tot = np.logaddexp(tot, np.logaddexp.reduce(o)); n += 10_000_000; best = max(best, float(o.max()))
    lnz_sweep = tot - np.log(n)
    print("nested ln Z", res["log_evidence"], "+-", res["log_evidence_err"], "sweep ln Z", lnz_sweep,
          "max logL nested", res["bestfit_params"]["log_likelihood"], "sweep", best, "ncall", res["num_likelihood_evaluations"])
    assert res["bestfit_params"]["log_likelihood"] >= best - 1.0
    assert abs(res["log_evidence"] - lnz_sweep) < 4 * res["log_evidence_err"] + 0.5
