"""``lightcurve-analysis`` on the GPU likelihood.

Mirrors the set-up half of the reference driver -- ``nmma/em/analysis.py:110-173`` (``analysis_setup``: filters, data,
time cuts, detection limits, light-curve model, systematics handler, priors, likelihood) with the option names of
``nmma/em/em_parsing.py`` -- and replaces the ``bilby.run_sampler`` half (``nmma/em/analysis.py:183-260``), whose
samplers evaluate one point per call and are not available offline, by the batched nested sampler of
``nmma_b200/samplers.py`` driven through ``EMTransientLikelihood.vectorized()``.  Outputs follow the reference's
naming: ``{outdir}/{label}_result.json`` and ``{outdir}/{label}_posterior_samples.dat``.

    python -m nmma_b200.em.analysis --model Bu2019lm --svd-path svdmodels --interpolation-type tensorflow \\
        --light-curve-data example_files/lightcurves/AT2017gfo.dat --trigger-time 57982.5285236896 \\
        --prior priors/Bu2019lm.prior --tmin 0.1 --tmax 14 --dt 0.5 --em-error-budget 1 --nlive 1024 \\
        --outdir outdir --label AT2017gfo
"""
from __future__ import annotations

import argparse
import json
import os
import time
from typing import Optional

import numpy as np

from ..core.priors import PriorDict, fixed_value, is_fixed_prior
from ..samplers import equal_weight, nested_sample
from . import io, utils
from .em_likelihood import EMTransientLikelihood
from .model import create_light_curve_model_from_args
from .prior import create_prior_from_args
from .systematics import FilterSystematicsHandler

__all__ = ["get_parser", "analysis_setup", "analysis", "main"]


def get_parser() -> argparse.ArgumentParser:
    """The subset of ``nmma/em/em_parsing.py`` that reaches the kilonova likelihood (same flags and dests)."""
    p = argparse.ArgumentParser(description="Inference on kilonova light curves with the B200 likelihood.")
    p.add_argument("--em-model", "--kilonova-model", "--model", dest="em_model", type=str, required=True)
    p.add_argument("--interpolation-type", type=str, default="tensorflow", help="tensorflow | sklearn_gp")
    p.add_argument("--svd-path", type=str, default="svdmodels")
    p.add_argument("--svd-mag-ncoeff", "--svd-ncoeff", dest="svd_mag_ncoeff", type=int, default=10)
    p.add_argument("--model-parameters", type=str, default=None)
    p.add_argument("--local-only", "--local-model-only", dest="local_only", action="store_true", default=True)
    p.add_argument("--outdir", type=str, default="outdir")
    p.add_argument("--label", type=str, required=True)
    p.add_argument("--trigger-time", type=float, default=None, help="MJD; default: first data point")
    p.add_argument("--light-curve-data", "--data", dest="light_curve_data", type=str, required=True)
    p.add_argument("--data-time-unit", type=str, default=None)
    p.add_argument("--prior", type=str, required=True)
    p.add_argument("--em-tmin", "--kilonova-tmin", "--tmin", dest="em_tmin", type=float, default=None)
    p.add_argument("--em-tmax", "--kilonova-tmax", "--tmax", dest="em_tmax", type=float, default=None)
    p.add_argument("--em-tstep", "--kilonova-tstep", "--dt", dest="em_tstep", type=float, default=None)
    p.add_argument("--data-tmin", type=float, default=0.0)
    p.add_argument("--data-tmax", type=float, default=np.inf)
    p.add_argument("--filters", type=str, default=None, help="comma separated; default: all filters in the data")
    p.add_argument("--em-error-budget", "--error-budget", "--kilonova-error", dest="em_error_budget", type=float,
                   default=1.0)
    p.add_argument("--systematics-file", type=str, default=None)
    p.add_argument("--detection-limit", type=float, default=None)
    p.add_argument("--remove-nondetections", action="store_true")
    p.add_argument("--use-Ebv", dest="use_Ebv", action="store_true", help="sample the extinction E(B-V)")
    p.add_argument("--Ebv-max", dest="Ebv_max", type=float, default=0.5724)
    p.add_argument("--em-extinction-law", dest="em_extinction_law", type=str, default=None,
                   help="P92_SMC_host (default) | G23_MW")
    p.add_argument("--verbose", action="store_true")
    # sampler (replaces bilby's --sampler / --nlive / --seed block)
    p.add_argument("--nlive", type=int, default=1024)
    p.add_argument("--dlogz", type=float, default=0.1)
    p.add_argument("--batch", type=int, default=16384, help="candidate points per likelihood launch")
    p.add_argument("--seed", "--sampler-seed", dest="seed", type=int, default=42)
    p.add_argument("--nposterior", type=int, default=5000, help="equal-weight posterior samples to write")
    p.add_argument("--max-calls", type=int, default=200_000_000, help="likelihood-evaluation budget of the sampler")
    return p


def analysis_setup(args, svd_mag_model=None):
    """``nmma/em/analysis.py:110-173`` for observed data: returns ``(priors, likelihood)``.

    `svd_mag_model`: an already loaded surrogate dict (tests, synthetic weights) instead of ``--svd-path`` files."""
    filters = utils.set_filters(args)
    data = io.load_em_observations(args, format="observations")
    trigger_time = args.trigger_time
    if trigger_time is None:  # read_trigger_time: fall back to the first observation
        trigger_time = min(float(np.min(d["time"])) for d in data.values())
    if filters is not None:
        data = {f: d for f, d in data.items() if f in filters}
    data = utils.cut_data_to_time_range(data, args, trigger_time, tmin=getattr(args, "data_tmin", 0.0),
                                        tmax=getattr(args, "data_tmax", np.inf))
    if getattr(args, "remove_nondetections", False):   # check_detections (em/utils.py:255-283)
        data = {f: {k: v[np.isfinite(d["mag_error"])] for k, v in d.items()} for f, d in data.items()}
    data = {f: d for f, d in data.items() if len(d["time"]) > 0}
    filters_to_analyze = [f for f in (filters or list(data)) if f in data]
    if not filters_to_analyze:
        raise ValueError("no observations left after the filter / time selection")
    detection_limit = utils.create_detection_limit(args, filters_to_analyze)

    if svd_mag_model is not None:
        from .model import SVDLightCurveModel
        light_curve_model = SVDLightCurveModel(args.em_model, svd_mag_model=svd_mag_model,
                                               interpolation_type=args.interpolation_type, filters=filters_to_analyze,
                                               sample_times=utils.setup_sample_times(args))
    else:
        light_curve_model = create_light_curve_model_from_args(args.em_model, args, filters=filters_to_analyze)
    light_curve_data = utils.setup_filtered_lc_data(data, trigger_time)
    handler = FilterSystematicsHandler(filters_to_analyze, args.systematics_file, args.em_error_budget, light_curve_data[0])
    priors = create_prior_from_args(args, handler)     # prior file + Ebv prior + em_syserr* priors (nmma/em/prior.py:221-244)
    light_curve_data = utils.check_model_time_consistency(light_curve_data, light_curve_model, priors, None)
    handler = FilterSystematicsHandler(filters_to_analyze, args.systematics_file, args.em_error_budget, light_curve_data[0])
    likelihood = EMTransientLikelihood(light_curve_model, light_curve_data, handler, priors, filters=filters_to_analyze,
                                       detection_limit=detection_limit)
    return priors, likelihood


def analysis(args, svd_mag_model=None) -> dict:
    """Set up, sample, write the result files; returns the result dictionary."""
    t0 = time.time()
    priors, likelihood = analysis_setup(args, svd_mag_model=svd_mag_model)
    columns = likelihood.columns
    transform, loglike = likelihood.vectorized(columns)

    def loglike_u(u):
        return np.asarray(loglike(transform(np.ascontiguousarray(u, dtype=np.float64))), dtype=float)

    loglike_u(np.full((1, len(columns)), 0.5))   # CUDA context, table upload, first launches: set-up, not sampling
    t1 = time.time()
    res = nested_sample(loglike_u, len(columns), nlive=args.nlive, batch=args.batch, dlogz=args.dlogz, seed=args.seed,
                        max_calls=args.max_calls)
    t2 = time.time()
    idx = equal_weight(res, args.nposterior, seed=args.seed)
    theta = np.asarray(transform(np.ascontiguousarray(res["samples_u"][idx])), dtype=float)
    logl = np.asarray(res["log_likelihoods"])[idx]
    best = int(np.argmax(res["log_likelihoods"]))
    best_theta = np.asarray(transform(np.ascontiguousarray(res["samples_u"][best:best + 1])), dtype=float)[0]
    fixed = {k: float(fixed_value(priors[k])) for k in priors if is_fixed_prior(priors[k])}
    posterior = {c: theta[:, i].tolist() for i, c in enumerate(columns)}
    posterior["log_likelihood"] = logl.tolist()
    result = {
        "label": args.label,
        "outdir": args.outdir,
        "sampler": "nmma_b200.samplers.nested_sample (single ellipsoid, batched)",
        "search_parameter_keys": list(columns),
        "fixed_parameter_keys": sorted(fixed),
        "log_evidence": res["log_evidence"],
        "log_evidence_err": res["log_evidence_err"],
        "log_noise_evidence": 0.0,
        "log_bayes_factor": res["log_evidence"],
        "information_gain": res["information"],
        "num_likelihood_evaluations": res["ncall"],
        "sampling_time": t2 - t1,
        "setup_time": t1 - t0,
        "nlive": args.nlive,
        "bestfit_params": dict(zip(columns, best_theta.tolist()), **fixed,
                               log_likelihood=float(res["log_likelihoods"][best])),
        "posterior": posterior,
    }
    os.makedirs(args.outdir, exist_ok=True)
    with open(os.path.join(args.outdir, f"{args.label}_result.json"), "w") as fh:
        json.dump(result, fh)
    header = " ".join(list(columns) + ["log_likelihood"])
    np.savetxt(os.path.join(args.outdir, f"{args.label}_posterior_samples.dat"), np.column_stack([theta, logl]),
               header=header, comments="")
    if args.verbose:
        print(f"ln Z = {res['log_evidence']:.3f} +- {res['log_evidence_err']:.3f}; {res['ncall']} likelihood "
              f"evaluations in {t2 - t1:.2f} s ({res['ncall'] / max(t2 - t1, 1e-9):.3g} evals/s incl. the sampler)")
    return result


def main(argv: Optional[list] = None):
    args = get_parser().parse_args(argv)
    return analysis(args)


if __name__ == "__main__":
    main()
