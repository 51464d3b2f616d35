"""Seeded random configurations through every kernel path the engine offers for them, against the oracle.

Each case draws: surrogate kind (MLP / GP), n_coeff, number of filters, observation tables (ragged counts, epochs outside the
model window, upper limits, one filter without data), sample grid (training grid / coarser uniform / geometric), constant
budget or YAML systematics, detection limits, a fixed-vs-sampled split of the priors, and a batch size that lands on a
different automatic path (latency path, filter-split, un-split tensor-core kernel, coefficient mode, fused GP kernel).
The point is the interplay: the individual features have their own tests in test_gpu_parity.py / test_gpu_round2.py."""
import copy

import numpy as np
import pytest

from helpers import SENTINEL, assert_logl_close, build_pair, synthetic_observations

pytestmark = pytest.mark.gpu

_YAML_TIME = {"config": {"withTime": {"value": True, "filters": [None], "time_nodes": 3, "type": "Uniform", "minimum": 0.1,
                                      "maximum": 1.5},
                         "withoutTime": {"value": False, "type": "Uniform", "minimum": 0, "maximum": 2}}}
_YAML_ONE = {"config": {"withTime": {"value": False, "filters": [None], "time_nodes": 4, "type": "Uniform", "minimum": 0,
                                     "maximum": 2},
                        "withoutTime": {"value": True, "type": "Uniform", "minimum": 0.05, "maximum": 1.5}}}
_ALL_FILTERS = ["ps1::g", "ps1::r", "ps1::i", "ps1::z", "sdssu", "2massj", "2massks", "ztfr"]


def _all_paths(eng, pts):
    out = {}
    for key, path, extra in (("two_stage", 2, None), ("fused_ffma", 1, "fused_supported"), ("tensor_core", 3, "tc_supported"),
                             ("fused_gp", 4, "gp_fused_supported"), ("latency", 5, "tc_front_supported")):
        if extra is not None and not eng.get_info(extra):
            continue
        eng.set_option("path", path)
        out[key] = eng.logl_host(pts)
    eng.set_option("path", 0)
    out["auto"] = eng.logl_host(pts)
    return out


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration(seed):
    import torch
    assert torch.cuda.is_available()
    from oracle import harness
    from nmma_b200 import synthetic as syn
    from nmma_b200.core.priors import Uniform
    from nmma_b200.mlmodel import random_surrogate
    rng = np.random.default_rng(1000 + seed)
    kind = "gp" if seed % 3 == 2 else "mlp"
    K = 10 if (kind == "gp" or rng.random() < 0.5) else int(rng.integers(4, 15))
    nf = int(rng.integers(1, 6))
    filters = list(rng.choice(_ALL_FILTERS, size=nf, replace=False))
    name = "Ka2017" if kind == "gp" else "Bu2019lm"
    mins, maxs = syn.GRID_BOUNDS[name]
    core = random_surrogate(filters, d=len(mins), kind=kind, seed=seed, K=K, H=int(rng.choice([256, 640, 2048])),
                            Ntr=int(rng.integers(40, 200)), param_mins=mins, param_maxs=maxs)
    counts = {f: int(rng.integers(1, 14)) for f in filters}
    obs_filters = list(filters)
    if nf > 2 and rng.random() < 0.5:
        obs_filters = obs_filters[:-1]                  # a model filter without data
    lc_data = synthetic_observations(obs_filters, rng, n_per_filter={f: counts[f] for f in obs_filters},
                                     tmin=0.2, tmax=float(rng.choice([9.0, 13.0, 13.0, 13.0, 22.5])),
                                     n_ul=int(rng.integers(0, 3)), mag0=float(rng.uniform(17, 19)), slope=float(rng.uniform(0.1, 0.3)))
    grid_kind = rng.choice(["training", "uniform", "geom"])
    sample_times = {"training": None, "uniform": np.arange(0.1, 20.0, 0.4), "geom": np.geomspace(0.05, 20.0, 60)}[grid_kind]
    sysk = rng.choice(["budget", "time", "one"])
    yaml = {"budget": None, "time": copy.deepcopy(_YAML_TIME), "one": copy.deepcopy(_YAML_ONE)}[sysk]
    u = rng.random()                                    # no limit / fainter than the data / (rarely) brighter: every point fails
    limit = np.inf if u < 0.5 else {f: float(rng.uniform(24.5, 27.0) if u < 0.9 else rng.uniform(20.0, 22.0)) for f in obs_filters}
    priors = syn.ka2017_prior() if kind == "gp" else syn.bu2019lm_prior()
    priors["luminosity_distance"] = Uniform(10.0, 120.0, name="luminosity_distance")
    priors["timeshift"] = Uniform(-1.0, 0.15, name="timeshift")
    if rng.random() < 0.4:
        key = "log10_vej" if kind == "gp" else "KNphi"
        priors[key] = float(rng.uniform(priors[key].minimum, priors[key].maximum))      # a fixed model parameter
    extras = []
    if rng.random() < 0.3:                              # redshift sampled independently of the distance (em/model.py:288-303)
        priors["redshift"] = Uniform(0.0, 0.08, name="redshift")
        extras.append("z")
    if kind == "mlp" and rng.random() < 0.3:            # a Constraint on a derived key (core/base.py:67-68)
        from nmma_b200.core.priors import Constraint
        priors["KNtheta"] = Constraint(minimum=10.0, maximum=80.0, name="KNtheta")
        extras.append("constraint")
    law, coef = None, None
    if rng.random() < 0.35:                             # dust: Ebv sampled, P92 SMC in the host frame or a linear law
        from types import SimpleNamespace
        from nmma_b200.em.prior import extinction_prior
        priors = extinction_prior(priors, SimpleNamespace(use_Ebv=True, Ebv_max=0.5724))
        law = str(rng.choice(["P92_SMC_host", "G23_MW"]))
        if law == "G23_MW":
            coef = {f: float(rng.uniform(0.3, 4.5)) for f in filters}
        extras.append(law)
    lik, olik, fixed, cols = build_pair(core, name, filters, obs_filters, lc_data, priors, kind=kind,
                                        sample_times=sample_times, error_budget=float(rng.uniform(0.3, 1.2)),
                                        systematics=yaml, detection_limit=limit, extinction_law=law, extinction_coef=coef)
    n = int(rng.choice([1, 37, 300, 1500, 5000])) if kind == "mlp" else int(rng.choice([5, 200, 4500]))
    pts, _ = priors.sample_array(n, np.random.default_rng(seed), cols)
    if n > 20:
        pts[3, 0] = np.nan
        pts[7, cols.index("timeshift")] = 30.0          # every detection outside the model window
    n_ref = min(n, 150)
    ref = harness.oracle_logl(olik, fixed, pts[:n_ref], cols)
    eng = lik.sub_model.engine_for(cols)
    results = _all_paths(eng, pts)
    tag = (f"seed {seed}: {kind} K={K} F={nf} obs={sum(counts[f] for f in obs_filters)} grid={grid_kind} sys={sysk} "
           f"limit={'inf' if limit is np.inf else 'finite'} {'+'.join(extras)} N={n} auto path {eng.get_info('last_path')}")
    for k, v in results.items():                        # diagnostics first: where the worst disagreement sits
        d = np.abs(v[:n_ref] - ref) / np.maximum(1.0, np.abs(ref))
        d[(v[:n_ref] == SENTINEL) | (ref == SENTINEL)] = 0.0
        i = int(np.argmax(d))
        if d[i] > 1e-4 or not np.array_equal(v[:n_ref] == SENTINEL, ref == SENTINEL):
            print(tag, "PATH", k, "worst row", i, "got", v[i], "ref", ref[i], "row", dict(zip(cols, pts[i])))
    worst = {k: assert_logl_close(v[:n_ref], ref) for k, v in results.items()}
    print("\n" + tag, {k: f"{e:.1e}" for k, e in worst.items()}, "sentinels", int((ref == SENTINEL).sum()), "of", n_ref)
    base = results["two_stage"]
    for k, v in results.items():                        # beyond the oracle sample: every path against the two-stage kernels
        assert np.array_equal(v == SENTINEL, base == SENTINEL), (tag, k)
        assert_logl_close(v, base)


def test_batch_size_sequence_on_one_engine():
    """One handle, batches of very different sizes back to back: every automatic path in turn, growing and shrinking
    scratch buffers (coefficients, partial sums, GP tickets), results independent of what ran before."""
    import torch
    assert torch.cuda.is_available()
    from nmma_b200 import synthetic as syn
    for kind, name, sizes in (("mlp", "Bu2019lm", [1, 5000, 37, 300, 70000, 2, 1500, 1024, 1025]),
                              ("gp", "Ka2017", [3, 6000, 200, 4096, 4095, 9000, 1])):
        filters = ["ps1::g", "ps1::r", "2massks", "sdssu"]
        core = syn.random_model(name, filters, kind=kind, seed=3, **({"Ntr": 120} if kind == "gp" else {}))
        rng = np.random.default_rng(5)
        lc_data = synthetic_observations(filters, rng, n_per_filter=9, tmax=12.0, n_ul=1, mag0=18.0, slope=0.2)
        priors = syn.ka2017_prior() if kind == "gp" else syn.bu2019lm_prior()
        lik, _, _, cols = build_pair(core, name, filters, filters, lc_data, priors, kind=kind)
        eng = lik.sub_model.engine_for(cols)
        big, _ = priors.sample_array(max(sizes), np.random.default_rng(9), cols)
        eng.set_option("path", 2)
        base = eng.logl_host(big[:6000])                 # plain two-stage kernels
        eng.set_option("path", 0)
        seen = set()
        for n in sizes + sizes[::-1]:
            got = eng.logl_host(big[:n])
            seen.add(eng.get_info("last_path"))
            m = min(n, 6000)
            assert_logl_close(got[:m], base[:m])
        print(kind, "automatic paths seen", sorted(seen))
        assert seen >= ({5, 3} if kind == "mlp" else {2, 4})
