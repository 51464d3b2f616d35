"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances are the north star's: magnitudes within 1e-3 mag, log L within
1e-4 * max(1, |log L|), failure (sentinel) masks and time indexing bit-exact.
"""
import json
import os

import numpy as np
import pytest

from helpers import (GOLDEN, SENTINEL, assert_logl_close, build_pair, fixture_core, oracle_ready_core,
                     synthetic_observations)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _paths(lik, pts, cols):
    """logL through every kernel path the configuration supports."""
    sub = lik.sub_model
    eng = sub.engine_for(cols)
    out = {}
    eng.set_option("path", 2)
    out["two_stage"] = eng.logl_host(pts)
    if eng.get_info("fused_supported"):
        eng.set_option("path", 1)
        for pt in (1, 2, 4):
            eng.set_option("points_per_thread", pt)
            out[f"fused_pt{pt}"] = eng.logl_host(pts)
            eng.set_option("no_fast_backend", 1)          # generic back end (two np.interp stages)
            out[f"fused_pt{pt}_generic"] = eng.logl_host(pts)
            eng.set_option("no_fast_backend", 0)
        eng.set_option("points_per_thread", 0)
    if eng.get_info("tc_supported"):
        eng.set_option("path", 3)                             # tcgen05 kernel (fp16 hi/lo split operands)
        out["fused_tc"] = eng.logl_host(pts)                  # small batches: filters split over CTAs (launch_tc.cu)
        eng.set_option("no_filter_split", 1)
        out["fused_tc_unsplit"] = eng.logl_host(pts)          # the throughput instantiation on the same points
        eng.set_option("no_filter_split", 0)
        eng.set_option("no_fast_backend", 1)
        out["fused_tc_generic"] = eng.logl_host(pts)
        eng.set_option("no_fast_backend", 0)
    if eng.get_info("tc_front_supported"):
        eng.set_option("path", 5)                             # latency path: filters and hidden ranges over all SMs + back end
        out["latency_hsplit"] = eng.logl_host(pts)
    if eng.get_info("gp_fused_supported"):
        eng.set_option("path", 4)                             # fused GP kernel (thread = point, item = (tile, filter))
        out["fused_gp"] = eng.logl_host(pts)
        eng.set_option("no_fast_backend", 1)
        out["fused_gp_generic"] = eng.logl_host(pts)
        eng.set_option("no_fast_backend", 0)
    eng.set_option("path", 0)
    return out


# ------------------------------------------------------------------------------------------------
# C1/C2: Bu2019lm-shaped tensorflow surrogate vs AT2017gfo
# ------------------------------------------------------------------------------------------------
def test_bu2019lm_at2017gfo_logl(torch_cuda):
    from nmma_b200 import synthetic as syn
    from oracle import harness
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    core = syn.random_model("Bu2019lm", filters, seed=0)
    priors = syn.bu2019lm_prior()
    lik, olik, fixed, cols = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors)
    pts, _ = priors.sample_array(700, np.random.default_rng(1234), cols)
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    assert np.isfinite(ref).all() and (ref != SENTINEL).sum() > 600
    for name, got in _paths(lik, pts, cols).items():
        err = assert_logl_close(got, ref)
        print(name, "max rel err", err)
    # dict entry point == batched entry point
    p0 = dict(zip(cols, pts[0]))
    assert lik.log_likelihood(p0) == pytest.approx(ref[0], rel=1e-4)


def test_bu2019lm_device_tensor_and_large_batch(torch_cuda):
    """Size-independent property at a larger N: permuting / tiling the points permutes log L exactly,
    and every kernel path returns identical values for identical rows."""
    torch = torch_cuda
    from nmma_b200 import synthetic as syn
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    core = syn.random_model("Bu2019lm", filters, seed=3)
    priors = syn.bu2019lm_prior()
    lik, olik, fixed, cols = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors)
    base, _ = priors.sample_array(4096, np.random.default_rng(7), cols)
    reps = 40                                              # 163,840 points: several full waves of tiles
    pts = np.tile(base, (reps, 1))
    perm = np.random.default_rng(0).permutation(len(pts))
    dev = torch.from_numpy(pts[perm]).cuda()
    out = lik.log_likelihood_batch(dev, cols)
    assert out.is_cuda and out.dtype == torch.float64
    out = out.cpu().numpy()
    unperm = np.empty_like(out)
    unperm[perm] = out
    unperm = unperm.reshape(reps, -1)
    assert np.array_equal(unperm, np.broadcast_to(unperm[0], unperm.shape)), "same row must give the same logL"
    small = lik.sub_model.engine_for(cols)
    small.set_option("path", 2)
    two = small.logl_host(base)
    small.set_option("path", 0)
    # the automatic path at this N is the tcgen05 kernel (fp16 hi/lo split operands, round-toward-zero accumulator): measured
    # 1.8e-6 relative against the fp32 two-stage kernels, 50x inside the 1e-4 north-star tolerance
    assert_logl_close(unperm[0], two, rtol=1e-5)
    assert small.get_info("tc_supported") == 1
    # the same rows through HOST buffers: the copy / compute pipeline (one-wave first block, whole-wave blocks, a last partial
    # block) must return what the one-launch device path returns, bit for bit, whatever block a row lands in
    host = lik.log_likelihood_batch(np.ascontiguousarray(pts[perm]), cols)
    assert isinstance(host, np.ndarray) and np.array_equal(host, out), "host pipeline and device path disagree"


# ------------------------------------------------------------------------------------------------
# Fixture weights (real trained surrogate): mags, golden value, logL
# ------------------------------------------------------------------------------------------------
def test_fixture_golden_magnitude(torch_cuda):
    """nmma/tests/joint_analysis_pipeline.py:108-120 through the GPU path: data['ztfr'][10] mag."""
    from nmma_b200.em import SVDLightCurveModel
    inj = json.load(open(os.path.join(GOLDEN, "bu2019lm_injection.json")))
    row = {k: v[0] for k, v in inj.items()}
    core = fixture_core("mlp", ("ztfr",))
    sample_times = np.arange(0.1, 10.0 + 0.5, 0.5)
    model = SVDLightCurveModel("Bu2019nsbh", svd_mag_model=core, interpolation_type="tensorflow",
                               filters=["ztfr"], sample_times=sample_times)
    params = model.parameter_conversion(dict(row))
    tobs, lc = model.gen_detector_lc(params)
    mask = tobs >= 0
    mags = lc["ztfr"][mask]
    noise = np.random.default_rng(42).normal(scale=0.1, size=len(mags))
    assert len(mags) == 18
    assert np.isclose((mags + noise)[10], 20.9294036584)
    assert abs((mags + noise)[10] - 20.9294036584) < 1e-5


@pytest.mark.parametrize("kind", ["mlp", "gp"])
def test_fixture_mags_and_logl(torch_cuda, kind):
    from oracle import harness, nmma_oracle as O
    from nmma_b200.core.priors import PriorDict, Sine, Uniform
    from nmma_b200.em import SVDLightCurveModel
    filters = ["ztfr", "sdssu", "2massks"]
    core = fixture_core(kind, filters)
    rng = np.random.default_rng(5)
    lc_data = synthetic_observations(filters, rng, n_per_filter=12, tmax=15.0, n_ul=2, mag0=19.0)
    priors = PriorDict()
    priors["luminosity_distance"] = Uniform(10.0, 200.0)
    priors["inclination_EM"] = Sine(0.0, np.pi / 2)
    priors["timeshift"] = Uniform(-0.2, 0.2)
    priors["log10_mej_dyn"] = Uniform(-2.2, -0.9)      # beyond the training range on purpose: no clipping
    priors["log10_mej_wind"] = Uniform(-2.2, -0.9)
    lik, olik, fixed, cols = build_pair(core, "Bu2019nsbh", filters, filters, lc_data, priors, kind=kind)
    n = 300 if kind == "mlp" else 60
    pts, _ = priors.sample_array(n, np.random.default_rng(11), cols)
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    for name, got in _paths(lik, pts, cols).items():
        print(kind, name, assert_logl_close(got, ref))
    # magnitudes: generate_lightcurve / gen_detector_lc vs the oracle, 1e-3 mag
    model = lik.sub_model.light_curve_model
    omodel = olik.light_curve_model
    worst = 0.0
    for row in pts[:20]:
        p = dict(fixed)
        p.update(dict(zip(cols, row)))
        p = model.parameter_conversion(p)
        tt = np.asarray(model.model_times, float)
        mine = model.generate_lightcurve(tt, dict(p))
        theirs = omodel.generate_lightcurve(tt, dict(p))
        t_m, app_m = model.gen_detector_lc(dict(p))
        t_o, app_o = omodel.gen_detector_lc(dict(p))
        assert np.array_equal(t_m, t_o), "detector-frame time grid must be bit-exact"
        for f in filters:
            assert np.array_equal(np.isfinite(mine[f]), np.isfinite(theirs[f]))
            worst = max(worst, np.abs(mine[f] - theirs[f]).max(), np.abs(app_m[f] - app_o[f]).max())
    assert worst < 1e-3
    print(kind, "max |dmag|", worst)


def test_mlp_fp32_accuracy_vs_fp64_truth(torch_cuda):
    """The fp32 summation order differs from NumPy's (and Keras'); both must sit equally close to
    the exact (fp64) network output."""
    torch = torch_cuda
    from nmma_b200.em import SVDLightCurveModel
    filters = ["ztfr", "sdssu", "2massks"]
    core = fixture_core("mlp", filters)
    model = SVDLightCurveModel("Bu2019nsbh", svd_mag_model=core, interpolation_type="tensorflow", filters=filters)
    eng = model._canonical_engine(np.asarray(model.model_times, float))
    rng = np.random.default_rng(2)
    x = rng.uniform([-2.0, -2.0, 0.0], [-1.05, -1.05, 90.0], size=(256, 3))
    pts = np.concatenate([x, np.full((256, 1), 40.0), np.zeros((256, 3))], axis=1)   # [x, dL, timeshift, redshift, Ebv]
    eng.set_option("path", 2)                      # plain two-stage kernels: fp32 FFMA front end (coeff_mlp_kernel)
    got = eng.coeffs(pts).cpu().numpy()
    eng.set_option("path", 0)                      # >= 128 points: the tensor-core kernel in coefficient mode (fp16 hi/lo split)
    got_tc = eng.coeffs(pts).cpu().numpy()
    for fi, f in enumerate(filters):
        W1, b1, W2, b2 = core[f]["model"]
        xs = ((x - core[f]["param_mins"]) / (core[f]["param_maxs"] - core[f]["param_mins"])).astype(np.float32)
        exact = np.maximum(xs.astype(np.float64) @ W1.astype(np.float64) + b1, 0) @ W2.astype(np.float64) + b2
        npf32 = np.maximum(xs @ W1 + b1, np.float32(0)) @ W2 + b2
        e_gpu = np.abs(got[:, fi, :] - exact).max()
        e_np = np.abs(npf32 - exact).max()
        e_tc = np.abs(got_tc[:, fi, :] - exact).max()
        print(f, "gpu fp32 err", e_gpu, "numpy fp32 err", e_np, "tensor-core fp16 hi/lo split err", e_tc)
        assert e_gpu < 2e-5 and e_gpu < 4 * e_np + 1e-6
        # the split drops the lo x lo products and the tensor core accumulates with round-toward-zero (group partials are
        # added with RN on the CUDA cores): <= ~1e-5 absolute on these trained weights (coefficients of order 1-10), i.e.
        # < 1e-4 mag through (maxs - mins) VA -- a factor 10 inside the 1e-3 mag budget
        assert e_tc < 3e-5


# ------------------------------------------------------------------------------------------------
# C3: Bu2023Ye-shaped (d = 7) + time/filter-dependent systematics + upper limits + detection limit
# ------------------------------------------------------------------------------------------------
SYS_YAMLS = {
    "legacy_with_time_all": {"config": {"withTime": {"value": True, "filters": [None], "time_nodes": 4,
                                                     "type": "Uniform", "minimum": 0, "maximum": 2},
                                        "withoutTime": {"value": False, "type": "Uniform", "minimum": 0, "maximum": 2}}},
    "legacy_without_time": {"config": {"withTime": {"value": False, "filters": [None], "time_nodes": 4,
                                                    "type": "Uniform", "minimum": 0, "maximum": 2},
                                       "withoutTime": {"value": True, "type": "Uniform", "minimum": 0, "maximum": 2}}},
    "legacy_groups": {"config": {"withTime": {"value": True, "filters": ["sdssu", ["2massj", "2massh"], "2massks",
                                                                         ["ps1::g", "ps1::r", "ps1::i", "ps1::z", "ps1::y"]],
                                              "time_nodes": 3, "type": "Uniform", "minimum": 0.1, "maximum": 2},
                                 "withoutTime": {"value": False, "type": "Uniform", "minimum": 0, "maximum": 2}}},
    "new_style_mixed": {"sdssu": {"prior": "Uniform(minimum=0.2, maximum=1.5)"},
                        "nir": {"filters": ["2massj", "2massh", "2massks"], "time_nodes": 3,
                                "prior": "Uniform(minimum=0.1, maximum=1.0)"},
                        "rest": {"time_range": "log 0.5 14 4", "prior": "Uniform(minimum=0.3, maximum=2.0)"}},
}


@pytest.mark.parametrize("sys_name", list(SYS_YAMLS))
@pytest.mark.parametrize("lim", [np.inf, 24.5])
def test_bu2023ye_systematics(torch_cuda, sys_name, lim):
    from nmma_b200 import synthetic as syn
    from oracle import harness
    import copy
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    core = syn.random_model("Bu2023Ye", filters, seed=1)
    priors = syn.bu2023ye_prior()
    priors["timeshift"].maximum = 0.1
    lik, olik, fixed, cols = build_pair(core, "Bu2023Ye", filters, filters, lc_data, priors,
                                        systematics=copy.deepcopy(SYS_YAMLS[sys_name]), detection_limit=lim)
    assert any(c.startswith("em_syserr") for c in cols)
    pts, _ = priors.sample_array(200, np.random.default_rng(99), cols)
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    for name, got in _paths(lik, pts, cols).items():
        print(sys_name, lim, name, assert_logl_close(got, ref), "sentinels", int((ref == SENTINEL).sum()))


# ------------------------------------------------------------------------------------------------
# C4: Ka2017-shaped sklearn GP across ZTF + PS1 filters with detection limits
# ------------------------------------------------------------------------------------------------
def test_ka2017_gp(torch_cuda):
    from nmma_b200 import synthetic as syn
    from oracle import harness
    filters = ["ztfg", "ztfr", "ztfi", "sdssu", "ps1::g", "ps1::r", "ps1::i", "ps1::z", "ps1::y"]
    core = syn.random_model("Ka2017", filters, kind="gp", seed=2, Ntr=329)
    rng = np.random.default_rng(8)
    lc_data = synthetic_observations(filters, rng, n_per_filter=10, tmax=13.0, n_ul=2, mag0=18.0, slope=0.15)
    limits = {"ztfg": 21.7, "ztfr": 21.4, "ztfi": 20.9, "sdssu": 23.9, "ps1::g": 25.0, "ps1::r": 24.7,
              "ps1::i": 24.0, "ps1::z": 23.3, "ps1::y": 22.1}
    priors = syn.ka2017_prior()
    priors["timeshift"].maximum = 0.2
    lik, olik, fixed, cols = build_pair(core, "Ka2017", filters, filters, lc_data, priors, kind="gp",
                                        detection_limit=limits)
    pts, _ = priors.sample_array(48, np.random.default_rng(21), cols)
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    got = lik.log_likelihood_batch(pts, cols)
    print("ka2017 gp", assert_logl_close(got, ref), "sentinels", int((ref == SENTINEL).sum()))
    for name, got_p in _paths(lik, pts, cols).items():
        print("ka2017 gp", name, assert_logl_close(got_p, ref))
    # fused GP kernel (gf_pow, per-tile tickets) against the two-stage kernels (rq_pow) on a batch with a ragged last tile:
    # with the generic fp64 back end both evaluate the same formulas, so they agree far below the tolerance
    eng = lik.sub_model.engine_for(cols)
    assert eng.get_info("gp_fused_supported") == 1
    big, _ = priors.sample_array(32 * 150 + 7, np.random.default_rng(22), cols)
    big[5, cols.index("log10_mej")] = np.nan
    eng.set_option("path", 2); two = eng.logl_host(big)
    eng.set_option("path", 4); eng.set_option("no_fast_backend", 1); gen = eng.logl_host(big)
    eng.set_option("no_fast_backend", 0); fast = eng.logl_host(big); again = eng.logl_host(big)
    eng.set_option("path", 0); auto = eng.logl_host(big)
    assert eng.get_info("last_path") == 4                     # the automatic path of a large GP batch
    sent = two == SENTINEL
    assert sent[5] and np.array_equal(sent, gen == SENTINEL) and np.array_equal(sent, fast == SENTINEL)
    ok = ~sent
    assert np.abs(gen[ok] - two[ok]).max() / np.maximum(1.0, np.abs(two[ok])).max() < 1e-9
    assert_logl_close(fast, two)
    assert np.array_equal(fast, again) and np.array_equal(fast, auto)   # schedule-independent sums, tickets reset
    # GP coefficients against sklearn.predict to 1e-9 relative (cancellation ~1e5 in k.alpha)
    eng = lik.sub_model.engine_for(cols)
    c_gpu = eng.coeffs(pts[:8]).cpu().numpy()
    for n in range(8):
        p = dict(fixed); p.update(dict(zip(cols, pts[n])))
        x = (np.array([p[k] for k in ["log10_mej", "log10_vej", "log10_Xlan"]]) - core[filters[0]]["param_mins"]) / \
            (core[filters[0]]["param_maxs"] - core[filters[0]]["param_mins"])
        for fi, f in enumerate(filters):
            c_ref = np.array([gp.predict(np.atleast_2d(x))[0] for gp in core[f]["gps"]])
            scale = np.abs(c_ref).max() + 1e-30
            assert np.abs(c_gpu[n, fi] - c_ref).max() / scale < 1e-7


# ------------------------------------------------------------------------------------------------
# Two-stage interpolation, averaged filters, edge cases
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("grid", ["tstep", "geom", "partial"])
def test_two_stage_sample_grid(torch_cuda, grid):
    """--em-tmin/--em-tmax grids: tt -> sample grid -> observation times."""
    from oracle import harness
    from nmma_b200 import synthetic as syn
    filters = ["ztfr", "sdssu", "2massks"]
    core = fixture_core("mlp", filters)
    if grid == "tstep":
        st = np.arange(0.1, 10.0 + 0.5, 0.5)
    elif grid == "geom":
        st = np.geomspace(0.05, 20.0, 150)
    else:
        st = np.linspace(-1.0, 25.0, 60)            # nodes outside [0, 21] are dropped in stage 2
    rng = np.random.default_rng(3)
    lc_data = synthetic_observations(filters, rng, n_per_filter=9, tmin=0.5, tmax=9.0, n_ul=1, mag0=19.0)
    from nmma_b200.core.priors import PriorDict, Sine, Uniform
    priors = PriorDict()
    priors["luminosity_distance"] = Uniform(20.0, 100.0)
    priors["inclination_EM"] = Sine(0.0, np.pi / 2)
    priors["timeshift"] = Uniform(-0.3, 0.3)
    priors["log10_mej_dyn"] = Uniform(-2.0, -1.05)
    priors["log10_mej_wind"] = Uniform(-2.0, -1.05)
    lik, olik, fixed, cols = build_pair(core, "Bu2019nsbh", filters, filters, lc_data, priors, sample_times=st)
    pts, _ = priors.sample_array(256, np.random.default_rng(4), cols)
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    for name, got in _paths(lik, pts, cols).items():
        print(grid, name, assert_logl_close(got, ref), "sentinels", int((ref == SENTINEL).sum()))


def test_out_of_range_detections_and_limits_give_sentinel(torch_cuda):
    """A detection outside the model's detector-frame window -> truncnorm NaN -> sentinel for the point;
    an upper limit outside -> logsf = 0; m > detection limit -> -inf -> sentinel."""
    from oracle import harness
    from nmma_b200.core.priors import PriorDict, Uniform
    filters = ["ztfr", "sdssu"]
    core = fixture_core("mlp", filters)
    times = {"ztfr": np.array([0.5, 3.0, 20.5, 24.0]), "sdssu": np.array([1.0, 2.0, 30.0])}
    mags = {"ztfr": np.array([19.0, 20.0, 23.0, 23.0]), "sdssu": np.array([20.0, 21.0, 22.0])}
    errs = {"ztfr": np.array([0.1, 0.1, 0.2, np.inf]), "sdssu": np.array([0.1, np.inf, np.inf])}
    lc_data = (times, mags, errs, 0.0)
    priors = PriorDict()
    priors["luminosity_distance"] = Uniform(20.0, 60.0)
    priors["KNtheta"] = Uniform(0.0, 90.0)
    priors["timeshift"] = Uniform(-1.5, 1.5)          # shifts the 20.5 d detection in and out of range
    priors["log10_mej_dyn"] = Uniform(-2.0, -1.05)
    priors["log10_mej_wind"] = Uniform(-2.0, -1.05)
    for lim in (np.inf, {"ztfr": 19.5, "sdssu": 25.0}):
        lik, olik, fixed, cols = build_pair(core, "Bu2019nsbh", filters, filters, lc_data, priors, detection_limit=lim)
        pts, _ = priors.sample_array(300, np.random.default_rng(6), cols)
        ref = harness.oracle_logl(olik, fixed, pts, cols)
        nsent = int((ref == SENTINEL).sum())
        if lim is np.inf:
            assert 0 < nsent < len(ref)
        else:
            assert nsent == len(ref)              # 20.0 > 19.5 limit in ztfr -> -inf for every point
        for name, got in _paths(lik, pts, cols).items():
            assert_logl_close(got, ref)


def test_averaged_filters(torch_cuda):
    """Observed filters without a model counterpart are averaged from model filters
    (nmma/em/utils.py:549-584); only the two-stage kernels serve this configuration."""
    from oracle import harness, nmma_oracle as O
    from nmma_b200.core.priors import PriorDict, Uniform
    # model filters named like the bare bands the averaging rules use
    src = fixture_core("mlp", ("ztfr", "sdssu", "2massks"))
    core = {"g": src["sdssu"], "r": src["ztfr"], "i": src["2massks"]}
    obs_filters = ["g", "r", "i", "w", "o", "V"]
    rng = np.random.default_rng(12)
    lc_data = synthetic_observations(obs_filters, rng, n_per_filter=6, tmax=10.0, n_ul=1, mag0=19.0)
    priors = PriorDict()
    priors["luminosity_distance"] = Uniform(20.0, 100.0)
    priors["KNtheta"] = Uniform(0.0, 90.0)
    priors["log10_mej_dyn"] = Uniform(-2.0, -1.05)
    priors["log10_mej_wind"] = Uniform(-2.0, -1.05)
    lik, olik, fixed, cols = build_pair(core, "Bu2019nsbh", ["g", "r", "i"], obs_filters, lc_data, priors)
    olik.model_filter_mapping, olik.obs_average_mapping = O.get_filter_name_mapping(obs_filters, {"g", "r", "i"})
    pts, _ = priors.sample_array(128, np.random.default_rng(13), cols)
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    eng = lik.sub_model.engine_for(cols)
    assert eng.get_info("fused_supported") == 0
    got = lik.log_likelihood_batch(pts, cols)
    print("averaged", assert_logl_close(got, ref))
    # From 128 points the coefficients come from the tensor-core kernel in coefficient mode (tcgen05 front end + generic
    # back end, launch_tc.cu: launch_tc_coeff); "path" = 2 keeps the plain two-stage kernels for comparison
    assert eng.get_info("tc_front_supported") == 1
    big, _ = priors.sample_array(256 * 3 + 77, np.random.default_rng(14), cols)          # filter-split and ragged tile
    big[3, cols.index("KNtheta")] = np.nan
    auto = eng.logl_host(big)
    eng.set_option("path", 2); plain = eng.logl_host(big); eng.set_option("path", 0)
    assert plain[3] == SENTINEL and auto[3] == SENTINEL
    print("averaged, tensor-core front end vs plain two-stage", assert_logl_close(auto, plain))
    ref_big = harness.oracle_logl(olik, fixed, big[:64], cols)
    assert_logl_close(auto[:64], ref_big)
    huge, _ = priors.sample_array(40_000, np.random.default_rng(15), cols)               # un-split instantiation
    a2 = eng.logl_host(huge)
    eng.set_option("path", 2); p2 = eng.logl_host(huge[:2000]); eng.set_option("path", 0)
    assert_logl_close(a2[:2000], p2)
    c_tc = eng.coeffs(huge[:4096]).cpu().numpy()
    eng.set_option("path", 2); c_plain = eng.coeffs(huge[:4096]).cpu().numpy(); eng.set_option("path", 0)
    assert np.abs(c_tc - c_plain).max() < 2e-5 * max(1.0, np.abs(c_plain).max())


def test_n_coeff_other_than_ten(torch_cuda):
    """n_coeff = 7: no fused instantiation; coefficients from the tensor-core kernel in coefficient mode (any K <= 16)."""
    from oracle import harness
    from nmma_b200 import synthetic as syn
    from nmma_b200.mlmodel import random_surrogate
    filters = ["ps1::g", "ps1::r", "2massks"]
    mins, maxs = syn.GRID_BOUNDS["Bu2019lm"]
    core = random_surrogate(filters, d=4, kind="mlp", seed=5, K=7, H=512, param_mins=mins, param_maxs=maxs)
    rng = np.random.default_rng(3)
    lc_data = synthetic_observations(filters, rng, n_per_filter=9, tmax=12.0, n_ul=1, mag0=18.0, slope=0.3)
    priors = syn.bu2019lm_prior()
    lik, olik, fixed, cols = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors)
    eng = lik.sub_model.engine_for(cols)
    assert eng.get_info("fused_supported") == 0 and eng.get_info("tc_supported") == 0 and eng.get_info("tc_front_supported") == 1
    pts, _ = priors.sample_array(1000, np.random.default_rng(4), cols)
    ref = harness.oracle_logl(olik, fixed, pts[:100], cols)
    got = eng.logl_host(pts)
    assert_logl_close(got[:100], ref)
    eng.set_option("path", 2); plain = eng.logl_host(pts); eng.set_option("path", 0)
    print("n_coeff 7", assert_logl_close(got, plain))


def test_obs_term_edge_semantics(torch_cuda):
    """SciPy wrapper semantics of SURVEY.md A.4, evaluated by the device function itself."""
    from scipy.stats import norm, truncnorm
    from nmma_b200.engine import KilonovaEngine
    eng = KilonovaEngine(0)
    inf, nan = np.inf, np.nan
    cases = [  # m, mu, sigma_obs, sigma_sys, lim
        (17.4, 17.5, 0.6, 0.8, inf), (18.0, inf, 0.6, 0.8, inf), (17.4, 17.5, 0.6, 0.8, 22.0),
        (23.0, 17.5, 0.6, 0.8, 22.0), (19.6, inf, inf, 1.0, inf), (19.6, 18.0, inf, 1.0, inf),
        (19.6, 0.0, inf, 0.5, inf), (19.6, 45.0, inf, 0.5, inf), (19.0, 18.0, 0.1, 0.0, inf),
        (19.0, 18.0, inf, 0.0, inf), (19.0, nan, 0.1, 1.0, inf), (19.0, 18.0, nan, 1.0, inf),
        (19.0, -inf, 0.1, 1.0, inf), (19.0, 18.0, 0.1, 1.0, 18.0), (19.0, 18.0, 0.1, 1.0, 19.0),
        (19.0, 30.0, 0.1, 1.0, 25.0), (19.0, 18.0, 0.1, -1.0, inf), (19.0, 18.0, inf, -1.0, inf),
    ]
    rng = np.random.default_rng(0)
    for _ in range(400):
        cases.append((rng.uniform(15, 25), rng.uniform(10, 40), rng.choice([rng.uniform(0.01, 0.5), inf]),
                      rng.uniform(0.05, 2.0), rng.choice([inf, rng.uniform(18, 26)])))
    m, mu, so, ss, lim = map(np.array, zip(*cases))
    got = eng.obs_terms(m, mu, so, ss, lim)
    ref = np.empty_like(got)
    with np.errstate(all="ignore"):
        for i in range(len(got)):
            sig = np.sqrt(so[i] ** 2 + ss[i] ** 2)
            if np.isfinite(sig):
                ref[i] = truncnorm.logpdf(m[i], -np.inf, (lim[i] - mu[i]) / sig, loc=mu[i], scale=sig)
            else:
                ref[i] = norm.logsf(m[i], mu[i], ss[i])
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref)
    assert np.array_equal(np.sign(got[~fin & ~np.isnan(ref)]), np.sign(ref[~fin & ~np.isnan(ref)]))
    rel = np.abs(got[fin] - ref[fin]) / np.maximum(1.0, np.abs(ref[fin]))
    assert rel.max() < 1e-12, rel.max()
    assert got[6] == pytest.approx(-772.9082649951413, rel=1e-13)


def test_nonfinite_inputs_and_empty_batch(torch_cuda):
    from nmma_b200 import synthetic as syn
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0, filters=["ps1::g", "sdssu"])
    core = syn.random_model("Bu2019lm", filters, seed=0)
    priors = syn.bu2019lm_prior()
    lik, olik, fixed, cols = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors)
    pts, _ = priors.sample_array(16, np.random.default_rng(3), cols)
    bad = pts.copy()
    bad[0, cols.index("log10_mej_dyn")] = np.nan
    bad[1, cols.index("luminosity_distance")] = -5.0
    bad[2, cols.index("timeshift")] = np.inf
    bad[3, cols.index("KNphi")] = np.inf
    for name, got in _paths(lik, bad, cols).items():
        assert (got[:4] == SENTINEL).all(), name
        assert (got[4:] != SENTINEL).all(), name
    assert lik.log_likelihood_batch(np.zeros((0, len(cols))), cols).shape == (0,)
    one = lik.log_likelihood_batch(pts[:1], cols)
    assert one.shape == (1,) and one[0] == lik.log_likelihood_batch(pts, cols)[0]


def test_abi_error_paths(torch_cuda):
    from nmma_b200 import _lib as L
    from nmma_b200.engine import KilonovaEngine
    eng = KilonovaEngine(0)
    with pytest.raises(L.NmmaB200Error) as ei:
        eng.P = 3
        eng.logl_host(np.zeros((2, 3)))
    assert ei.value.code == L.ERR_STATE
    with pytest.raises(L.NmmaB200Error):
        KilonovaEngine(10 ** 6)


def test_tensor_core_fp16_staging_dynamic_range(torch_cuda):
    """The tensor-core front end carries every fp32 operand as two fp16 values; what keeps them in the fp16 range are exact
    power-of-two scalings (rows of [W1; b1], columns of W2, one scale per point and filter -- csrc/tc_kernel.cuh).  Weights whose
    rows / columns differ by many orders of magnitude and points far outside the training range must come out as close to
    the exact (fp64) network as the fp32 FFMA kernel does."""
    torch = torch_cuda
    from nmma_b200.em import SVDLightCurveModel
    filters = ["ztfr", "sdssu", "2massks"]
    core = fixture_core("mlp", filters)
    rng = np.random.default_rng(11)
    row_scale = np.array([1e-5, 3e2, 1.0])             # layer-1 rows: 7.5 orders of magnitude apart
    col_scale = 10.0 ** rng.uniform(-6, 4, size=10)    # layer-2 columns: 10 orders
    for f in filters:
        W1, b1, W2, b2 = (np.array(a, dtype=np.float32) for a in core[f]["model"])
        W1 = (W1 * row_scale[:, None]).astype(np.float32)
        b1 = (b1 * np.float32(1e-3)).astype(np.float32)
        W2 = (W2 * col_scale[None, :]).astype(np.float32)
        W2[::7] *= np.float32(1e-9)                     # entries far below their column's maximum (fp16 subnormal remainders)
        b2 = (b2 * col_scale).astype(np.float32)
        core[f]["model"] = (W1, b1, W2, b2)
    model = SVDLightCurveModel("Bu2019nsbh", svd_mag_model=core, interpolation_type="tensorflow", filters=filters)
    eng = model._canonical_engine(np.asarray(model.model_times, float))
    lo, hi = core[filters[0]]["param_mins"], core[filters[0]]["param_maxs"]
    # scaled inputs from 1e-4 to 1e4 times the training range, both signs, and the exact corners 0 and 1
    mag = 10.0 ** rng.uniform(-4, 4, size=(512, 3)) * rng.choice([-1.0, 1.0], size=(512, 3))
    mag[:8] = rng.integers(0, 2, size=(8, 3))
    x = lo + mag * (hi - lo)
    pts = np.concatenate([x, np.full((512, 1), 40.0), np.zeros((512, 3))], axis=1)   # [x, dL, timeshift, redshift, Ebv]
    eng.set_option("path", 2)                      # plain two-stage kernels: fp32 FFMA front end (coeff_mlp_kernel)
    got = eng.coeffs(pts).cpu().numpy()
    eng.set_option("path", 0)                      # >= 128 points: the tensor-core kernel in coefficient mode
    got_tc = eng.coeffs(pts).cpu().numpy()
    assert np.isfinite(got_tc).all()
    for fi, f in enumerate(filters):
        W1, b1, W2, b2 = core[f]["model"]
        xs = ((x - core[f]["param_mins"]) / (core[f]["param_maxs"] - core[f]["param_mins"])).astype(np.float32)
        h = np.maximum(xs.astype(np.float64) @ W1.astype(np.float64) + b1, 0)
        exact = h @ W2.astype(np.float64) + b2
        # what the rounding errors are relative to: sum_i |x_i W1_ij| + |b1_j| per hidden unit (layer 1 may cancel), through |W2|
        h_abs = np.abs(xs.astype(np.float64)) @ np.abs(W1.astype(np.float64)) + np.abs(b1)
        scale = h_abs @ np.abs(W2.astype(np.float64)) + np.abs(b2)
        e_ffma = (np.abs(got[:, fi, :] - exact) / scale).max()
        e_tc = (np.abs(got_tc[:, fi, :] - exact) / scale).max()
        print(f, "relative to sum |x W1| |W2|: FFMA", e_ffma, "tensor core", e_tc)
        # measured: FFMA 3.5-4.3e-8, tensor core 1.9-2.4e-7 (22-bit operands, round-toward-zero accumulation in chains of 32 MMAs)
        assert e_ffma < 2e-7
        assert e_tc < 1e-6
