"""GPU parity tests added in round 2 (VERDICT r01 "parity holes"): the device transforms of
``observation_angle_conversion`` / the log10 twins, a sampled ``redshift``, the legacy ``OpticalLightCurve`` entry, the
dynesty pool seam, Constraint priors, extinction (Ebv != 0), and observation times on / one ulp around the
detector-frame grid nodes and range ends.  Same tolerances as tests/test_gpu_parity.py (north star)."""
import numpy as np
import pytest

from helpers import SENTINEL, assert_logl_close, build_pair, fixture_core, synthetic_observations
from test_gpu_parity import _paths

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


# ------------------------------------------------------------------------------------------------
# a2 / a3: every device transform (nmma/core/conversion.py:119-126, nmma/em/model.py:272-286) and Z_PARAM
# ------------------------------------------------------------------------------------------------
def _twin_priors(angle, sampled_redshift):
    """Bu2023Ye-shaped (d = 7) prior whose columns reach the model only through a conversion:
    mej_dyn -> log10_mej_dyn (XF_LOG10), log10_vej_dyn -> vej_dyn (XF_POW10), `angle` -> KNtheta."""
    from nmma_b200.core.priors import Cosine, PriorDict, Sine, Uniform
    p = PriorDict()
    p["mej_dyn"] = Uniform(10 ** -3.0, 10 ** -1.7, name="mej_dyn")
    p["log10_vej_dyn"] = Uniform(np.log10(0.12), np.log10(0.25), name="log10_vej_dyn")
    p["Yedyn"] = Uniform(0.15, 0.30)
    p["log10_mej_wind"] = Uniform(-2.0, -0.89)
    p["vej_wind"] = Uniform(0.03, 0.15)
    p["Yewind"] = Uniform(0.20, 0.40)
    if angle == "theta_jn":
        p["theta_jn"] = Sine(0.0, np.pi, name="theta_jn")                # folded to <= pi/2 on the device (XF_THETAJN_DEG)
    elif angle == "cos_theta_jn":
        p["cos_theta_jn"] = Uniform(-1.0, 1.0, name="cos_theta_jn")      # arccos, fold, degrees (XF_COSTHETAJN_DEG)
    elif angle == "inclination_EM":
        p["inclination_EM"] = Sine(0.0, np.pi / 2)
    else:
        p["KNtheta"] = Uniform(0.0, 90.0)
    if sampled_redshift:
        p["redshift"] = Uniform(0.002, 0.045, name="redshift")             # Z_PARAM: get_redshift returns the sampled value
        p["luminosity_distance"] = Uniform(10.0, 200.0)                    # independent of it, as the reference allows
    else:
        p["luminosity_distance"] = Uniform(10.0, 200.0)
    p["timeshift"] = Uniform(-2.0, 0.1)
    return p


@pytest.mark.parametrize("angle,sampled_redshift", [("theta_jn", False), ("cos_theta_jn", True), ("KNtheta", True),
                                                      ("inclination_EM", False)])
def test_device_transforms_and_sampled_redshift(torch_cuda, angle, sampled_redshift):
    from nmma_b200 import _lib as L
    from nmma_b200 import synthetic as syn
    from oracle import harness
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    core = syn.random_model("Bu2023Ye", filters, seed=4)
    priors = _twin_priors(angle, sampled_redshift)
    lik, olik, fixed, cols = build_pair(core, "Bu2023Ye", filters, filters, lc_data, priors)
    plan = lik.sub_model.plan_layout(cols)
    xf = [s.transform for s in plan["xsrc"]]
    assert xf[0] == L.XF_LOG10 and xf[1] == L.XF_POW10
    assert xf[6] == {"theta_jn": L.XF_THETAJN_DEG, "cos_theta_jn": L.XF_COSTHETAJN_DEG, "KNtheta": L.XF_NONE,
                     "inclination_EM": L.XF_RAD2DEG}[angle]
    assert plan["zmode"] == (L.Z_PARAM if sampled_redshift else L.Z_TABLE)
    pts, _ = priors.sample_array(400, np.random.default_rng(17), cols)
    if angle == "cos_theta_jn":     # exact ends of the arccos domain
        pts[0, cols.index("cos_theta_jn")] = 1.0
        pts[1, cols.index("cos_theta_jn")] = -1.0
        pts[2, cols.index("cos_theta_jn")] = 0.0
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    assert (ref != SENTINEL).sum() > 300
    for name, got in _paths(lik, pts, cols).items():
        print(angle, sampled_redshift, name, assert_logl_close(got, ref))
    # KNtheta seen by the surrogate: compare the device's scaled inputs through the magnitudes of one point
    p0 = dict(fixed); p0.update(dict(zip(cols, pts[5])))
    assert lik.log_likelihood(p0) == pytest.approx(ref[5], rel=1e-4)


# ------------------------------------------------------------------------------------------------
# g1: the north star's `OpticalLightCurve.log_likelihood`
# ------------------------------------------------------------------------------------------------
def test_optical_light_curve_dict_entry(torch_cuda):
    from nmma_b200 import synthetic as syn
    from nmma_b200.em import OpticalLightCurve, SVDLightCurveModel
    from oracle import harness
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    times, mags, errs, trig = lc_data
    core = syn.random_model("Bu2019lm", filters, seed=0)
    priors = syn.bu2019lm_prior()
    model = SVDLightCurveModel("Bu2019lm", svd_mag_model=core, interpolation_type="tensorflow", filters=filters)
    legacy = {f: np.c_[times[f] + trig, mags[f], errs[f]] for f in filters}       # absolute MJD rows [t, mag, err]
    lik = OpticalLightCurve(model, filters, legacy, trig, error_budget=1.0, tmin=0.0, tmax=14.0, priors=priors)
    olik, fixed = harness.build_oracle_likelihood(core, model.model_parameters, filters, np.asarray(model.model_times, float),
                                                  filters, lc_data, priors, z_table=model._z_table)
    cols = lik.columns
    pts, _ = priors.sample_array(24, np.random.default_rng(3), cols)
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    got = np.array([lik.log_likelihood(dict(zip(cols, row))) for row in pts])
    assert_logl_close(got, ref)
    assert isinstance(lik.log_likelihood(dict(zip(cols, pts[0]))), float)
    # a fixed key given a different value in the dict is evaluated as given, like the reference does
    priors2 = syn.bu2019lm_prior()
    priors2["timeshift"] = -0.5
    lik2 = OpticalLightCurve(model, filters, legacy, trig, error_budget=1.0, priors=priors2)
    p = dict(zip(cols, pts[1]))
    assert lik2.log_likelihood(p) == pytest.approx(ref[1], rel=1e-4)                # dict value wins over the prior constant
    p_fixed = dict(p, timeshift=-0.5)
    want = harness.oracle_logl(olik, fixed, np.array([[p_fixed[c] for c in cols]]), cols)[0]
    assert lik2.log_likelihood(p_fixed) == pytest.approx(want, rel=1e-4)
    # without priors every numeric key of the first call becomes a column
    model3 = SVDLightCurveModel("Bu2019lm", svd_mag_model=core, interpolation_type="tensorflow", filters=filters)
    lik3 = OpticalLightCurve(model3, filters, legacy, trig, error_budget=1.0)
    with pytest.raises(ValueError):          # luminosity_distance without a prior: the reference's per-call z_at_value path
        lik3.log_likelihood(dict(zip(cols, pts[0])))
    p3 = dict(zip(cols, pts[2]))
    z3 = float(np.interp(p3.pop("luminosity_distance"), *model._z_table))
    lik4 = OpticalLightCurve(SVDLightCurveModel("Bu2019lm", svd_mag_model=core, interpolation_type="tensorflow", filters=filters),
                             filters, legacy, trig, error_budget=1.0)
    p3["redshift"] = z3                      # dict with an explicit redshift and the default distance (1e-5 Mpc = 10 pc)
    want3 = harness.oracle_logl(harness.build_oracle_likelihood(core, model.model_parameters, filters,
                                                                np.asarray(model.model_times, float), filters, lc_data, {})[0],
                                {}, np.array([list(p3.values())]), list(p3.keys()))[0]
    assert lik4.log_likelihood(p3) == pytest.approx(want3, rel=1e-4)


# ------------------------------------------------------------------------------------------------
# 8b: the dynesty pool seam
# ------------------------------------------------------------------------------------------------
def test_pool_map_on_gpu(torch_cuda):
    from nmma_b200 import synthetic as syn
    from oracle import harness
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    core = syn.random_model("Bu2019lm", filters, seed=0)
    priors = syn.bu2019lm_prior()
    lik, olik, fixed, cols = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors)
    pool = lik.pool
    pts, _ = priors.sample_array(96, np.random.default_rng(2), cols)
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    eng = lik.sub_model.engine_for(cols)
    l0 = eng.get_info("launches")
    got = pool.map(pool.loglike, [row for row in pts])                      # array thetas
    assert eng.get_info("launches") - l0 <= 2, "one batched GPU call per map, not one per point"
    assert_logl_close(got, ref)
    got_d = pool.map(lik.log_likelihood, [dict(zip(cols, row)) for row in pts])    # dict thetas, bound method of the likelihood
    assert_logl_close(got_d, ref)

    class Wrapped:                                                           # dynesty's _function_wrapper shape
        def __init__(self, func):
            self.func = func

        def __call__(self, x):
            return self.func(x)

    assert_logl_close(pool.map(Wrapped(pool.loglike), list(pts)), ref)
    # prior transform through the same pool: unit cube in, physical points out (never log-likelihoods)
    u = np.random.default_rng(5).uniform(size=(64, len(cols)))
    phys = np.array(pool.map(pool.prior_transform, list(u)))
    want = np.stack(priors.rescale(cols, u), axis=1)
    assert np.allclose(phys, want, rtol=1e-13, atol=0)
    # a foreign function with P-long arrays is mapped verbatim
    assert pool.map(lambda t: float(np.sum(t)), list(u)) == [float(np.sum(t)) for t in u]
    assert pool.loglike(pts[3]) == pytest.approx(ref[3], rel=1e-4)


# ------------------------------------------------------------------------------------------------
# a1: Constraint priors in the batched entry (nmma/core/base.py:67-68)
# ------------------------------------------------------------------------------------------------
def test_constraint_priors_batched(torch_cuda):
    from nmma_b200 import synthetic as syn
    from nmma_b200.core.priors import Constraint
    from oracle import harness
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    core = syn.random_model("Bu2019lm", filters, seed=0)
    priors = syn.bu2019lm_prior()
    priors["KNtheta"] = Constraint(minimum=20.0, maximum=70.0, name="KNtheta")         # derived: inclination_EM * 180 / pi
    lik, olik, fixed, cols = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors)
    assert "KNtheta" not in cols and list(lik.constraints) == ["KNtheta"]
    pts, _ = priors.sample_array(600, np.random.default_rng(8), cols)
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    kn = np.degrees(pts[:, cols.index("inclination_EM")])
    assert np.array_equal(ref == SENTINEL, ~((kn > 20.0) & (kn < 70.0)))               # (no other failures in this configuration)
    for name, got in _paths(lik, pts, cols).items():
        assert_logl_close(got, ref)
    assert lik.log_likelihood(dict(zip(cols, pts[0]))) == pytest.approx(ref[0], rel=1e-4)
    # a constraint directly on a sampled column
    priors2 = syn.bu2019lm_prior()
    from collections import OrderedDict
    priors2["log10_mej_wind"] = Constraint(minimum=-2.0, maximum=-1.0)
    lik2, olik2, fixed2, cols2 = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors2)
    assert "log10_mej_wind" not in cols2
    with pytest.raises(AttributeError):                                                  # the model parameter has no source left
        lik2.log_likelihood_batch(pts[:4, :len(cols2)], cols2)
    assert isinstance(OrderedDict(), dict)


# ------------------------------------------------------------------------------------------------
# f4: extinction, Ebv != 0 (nmma/em/model.py:323-350, nmma/em/utils.py:373-433)
# ------------------------------------------------------------------------------------------------
def test_extinction_mags_all_filters(torch_cuda):
    """ext_mag per filter and point: gen_detector_lc with Ebv != 0 against the oracle on every filter of the
    Bu2019nsbh surrogate that has a wavelength entry (26 names; the five Bessell bands have no vendored table and
    stay uncorrected, as the reference leaves filters it cannot find)."""
    from nmma_b200 import synthetic as syn
    from nmma_b200.em import SVDLightCurveModel
    from oracle import nmma_oracle as O, harness
    names = ["2massh", "2massj", "2massks", "atlasc", "atlaso", "bessellb", "besselli", "bessellr", "bessellux", "bessellv",
             "ps1::g", "ps1::i", "ps1::r", "ps1::y", "ps1::z", "sdssu", "uvot::b", "uvot::u", "uvot::uvm2", "uvot::uvw1",
             "uvot::uvw2", "uvot::v", "uvot::white", "ztfg", "ztfi", "ztfr"]
    core = syn.random_model("Bu2019nsbh", names, seed=6)
    model = SVDLightCurveModel("Bu2019nsbh", svd_mag_model=core, interpolation_type="tensorflow", filters=names)
    assert len(model.default_filts) == 21 and not any(f.startswith("bessell") for f in model.default_filts)
    omodel = O.OracleSVDLightCurveModel(model.model_parameters, harness.oracle_core(core), filters=names,
                                        sample_times=np.asarray(model.model_times, float),
                                        default_filts=model.default_filts, lambdas=model.lambdas)
    rng = np.random.default_rng(1)
    worst = 0.0
    for ebv, z in [(0.0, 0.01), (0.05, 0.0), (0.3, 0.0098), (0.5724, 0.045), (1.5, 0.3), (0.2, 8.0)]:
        p = {"log10_mej_dyn": rng.uniform(-2, -1.1), "log10_mej_wind": rng.uniform(-2, -1.1), "KNtheta": rng.uniform(0, 90),
             "luminosity_distance": 40.0, "redshift": z, "timeshift": 0.1, "Ebv": ebv}
        t_m, app_m = model.gen_detector_lc(dict(p))
        t_o, app_o = omodel.gen_detector_lc(dict(p))
        abs_m = model.generate_lightcurve(np.asarray(model.model_times, float), dict(p))
        assert np.array_equal(t_m, t_o)
        for f in names:
            assert np.array_equal(np.isfinite(app_m[f]), np.isfinite(app_o[f]))
            fin = np.isfinite(app_o[f])
            worst = max(worst, np.abs(app_m[f][fin] - app_o[f][fin]).max())
            if f.startswith("bessell") or ebv == 0.0:      # no entry / no dust: apparent - absolute is the distance term only
                d = (app_m[f] - abs_m[f])[fin]
                assert np.ptp(d) < 1e-9
        if ebv > 0 and z < 1:
            shift = (app_m["uvot::uvw2"] - abs_m["uvot::uvw2"]) - (app_m["2massks"] - abs_m["2massks"])
            assert np.nanmin(shift) > 2.0 * ebv              # UV extinguished far more than K band (SMC far-UV rise)
    print("extinction: max |dmag| vs oracle", worst)
    assert worst < 1e-3


@pytest.mark.parametrize("law", ["P92_SMC_host", "G23_MW"])
def test_extinction_logl_sampled_ebv(torch_cuda, law):
    """C2-shaped run (Bu2019lm vs AT2017gfo) with Ebv sampled from the reference's triangular prior."""
    from types import SimpleNamespace
    from nmma_b200 import synthetic as syn
    from nmma_b200.em.prior import extinction_prior
    from oracle import harness
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    core = syn.random_model("Bu2019lm", filters, seed=0)
    priors = extinction_prior(syn.bu2019lm_prior(), SimpleNamespace(use_Ebv=True, Ebv_max=0.5724))
    coef = None
    if law == "G23_MW":   # the G23 curve is third-party data: any per-filter A_f / E(B-V) exercises the linear law
        coef = dict(zip(filters, [3.7, 2.7, 2.0, 1.5, 1.25, 0.8, 0.5, 0.35, 4.8]))
    lik, olik, fixed, cols = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors, extinction_law=law,
                                        extinction_coef=coef)
    assert cols[-1] == "Ebv"
    pts, _ = priors.sample_array(500, np.random.default_rng(77), cols)
    pts[0, -1] = 0.0                                        # Ebv == 0 exactly: the reference skips the correction
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    noext = pts.copy(); noext[:, -1] = 0.0
    ref0 = harness.oracle_logl(olik, fixed, noext, cols)
    assert np.abs(ref - ref0)[1:].max() > 1.0               # the dust matters in this configuration
    for name, got in _paths(lik, pts, cols).items():
        print(law, name, assert_logl_close(got, ref))
    # device prior for the Ebv column: sweep == batch on the same draws
    out, p = lik.log_likelihood_sweep(3000, seed=5, return_points=True, columns=cols)
    again = lik.log_likelihood_batch(p, cols)
    assert np.array_equal(out.cpu().numpy(), again.cpu().numpy())
    e = p[:, -1].cpu().numpy()
    assert e.min() >= 0 and e.max() <= 0.5724 and abs(e.mean() - 0.5724 / 3) < 0.02      # triangular density, mean = max / 3
    bad = pts[:3].copy(); bad[0, -1] = np.nan; bad[1, -1] = np.inf
    got = lik.log_likelihood_batch(bad, cols)
    assert got[0] == SENTINEL and got[1] == SENTINEL and got[2] != SENTINEL


# ------------------------------------------------------------------------------------------------
# interval indices / range masks: observation times on and one ulp around the detector-frame nodes
# ------------------------------------------------------------------------------------------------
def test_observations_on_grid_nodes_and_range_ends(torch_cuda):
    """np.interp semantics at the discrete decisions: t == node, t one ulp either side of a node, t == first / last node
    (in range) and one ulp outside (left = right = +inf -> detection gives the sentinel, upper limit gives log 1 = 0).
    With fixed z and timeshift every point shares the detector-frame grid t_j = fl(fl(s_j (1+z)) + ts), so the crafted
    times hit the nodes for EVERY point; all kernel paths must reproduce the oracle's sentinel mask bit for bit."""
    from nmma_b200.core.priors import PriorDict, Uniform
    from oracle import harness
    filters = ["ztfr", "sdssu"]
    core = fixture_core("mlp", filters)
    tt = core["ztfr"]["tt"]
    z, ts = 0.0115438, -0.3171
    tobs = tt * (1 + z) + ts                                  # the reference's expression, same two roundings
    up = lambda x: np.nextafter(x, np.inf)     # noqa: E731
    dn = lambda x: np.nextafter(x, -np.inf)    # noqa: E731
    inner = np.array([tobs[7], up(tobs[7]), dn(tobs[7]), tobs[100], up(tobs[100]), dn(tobs[100]), tobs[209], dn(tobs[210]),
                      tobs[0], up(tobs[0]), tobs[210], 0.5 * (tobs[33] + tobs[34])])
    cases = {"inside": (np.sort(inner), None),
             "below_first_ul": (np.r_[dn(tobs[0]), inner[:4]], 0),       # index of the out-of-range row -> made an upper limit
             "above_last_ul": (np.r_[inner[:4], up(tobs[210])], 4),
             "below_first_det": (np.r_[dn(tobs[0]), inner[:4]], None),
             "above_last_det": (np.r_[inner[:4], up(tobs[210])], None)}
    priors = PriorDict()
    priors["luminosity_distance"] = Uniform(20.0, 100.0)
    priors["redshift"] = z
    priors["timeshift"] = ts
    priors["KNtheta"] = Uniform(0.0, 90.0)
    priors["log10_mej_dyn"] = Uniform(-2.0, -1.05)
    priors["log10_mej_wind"] = Uniform(-2.0, -1.05)
    rng = np.random.default_rng(0)
    for name, (t, ul) in cases.items():
        order = np.argsort(t)
        t = t[order]
        err = np.full(len(t), 0.1)
        if ul is not None:
            err[np.nonzero(order == ul)[0][0]] = np.inf
        times = {"ztfr": t, "sdssu": np.array([1.0, 2.5, 4.0])}
        mags = {"ztfr": 19.0 + 0.2 * t + rng.normal(scale=0.2, size=len(t)), "sdssu": np.array([20.0, 21.0, 22.0])}
        errs = {"ztfr": err, "sdssu": np.array([0.1, 0.2, np.inf])}
        lik, olik, fixed, cols = build_pair(core, "Bu2019nsbh", filters, filters, (times, mags, errs, 0.0), priors)
        pts, _ = priors.sample_array(160, np.random.default_rng(9), cols)
        ref = harness.oracle_logl(olik, fixed, pts, cols)
        if name.endswith("_det"):
            assert (ref == SENTINEL).all(), name            # a detection outside the window: truncnorm NaN -> sentinel
        else:
            assert (ref != SENTINEL).all(), name
        for path, got in _paths(lik, pts, cols).items():
            assert_logl_close(got, ref)
    # sampled timeshift: for each point put one observation exactly on ITS node 57 -- the FAST back end's fp32 index guess
    # lands within fast_delta of an integer and must be settled by the exact comparisons
    priors["timeshift"] = Uniform(-0.4, 0.2)
    base_t = np.array([0.8, 2.0, 5.5, 9.0])
    times = {"ztfr": base_t, "sdssu": np.array([1.0, 2.5, 4.0])}
    mags = {"ztfr": 19.0 + 0.2 * base_t, "sdssu": np.array([20.0, 21.0, 22.0])}
    errs = {"ztfr": np.full(4, 0.1), "sdssu": np.array([0.1, 0.2, np.inf])}
    lik, olik, fixed, cols = build_pair(core, "Bu2019nsbh", filters, filters, (times, mags, errs, 0.0), priors)
    pts, _ = priors.sample_array(200, np.random.default_rng(10), cols)
    ic = cols.index("timeshift")
    hits = 0
    for i in range(len(pts)):       # choose the timeshift so that node j of this point equals an observation time exactly
        target = base_t[i % 4]
        j = int(round((target + 0.02 + 0.003 * (i % 50)) / (1 + z) / 0.1))
        cand = target - tt[j] * (1 + z)
        for c in (cand, np.nextafter(cand, np.inf), np.nextafter(cand, -np.inf)):
            if tt[j] * (1 + z) + c == target and -0.4 <= c <= 0.2:
                pts[i, ic] = c
                hits += 1
                break
    assert hits > 20
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    for path, got in _paths(lik, pts, cols).items():
        print("on-node", path, assert_logl_close(got, ref))


# ------------------------------------------------------------------------------------------------
# latency path: CUDA-graph replay per batch size, invalidation on reconfiguration
# ------------------------------------------------------------------------------------------------
def test_latency_graph_replay_and_invalidation(torch_cuda):
    """nmma_b200_logl_host, N <= 256: the first call of a size runs un-captured, the second captures a CUDA graph, later ones
    replay it.  Replays must follow the inputs, survive other batch sizes in between, and be dropped when the configuration
    (here: the detection limit) or an option changes."""
    from nmma_b200 import synthetic as syn
    from oracle import harness
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    core = syn.random_model("Bu2019lm", filters, seed=0)
    priors = syn.bu2019lm_prior()
    lik, olik, fixed, cols = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors)
    pts, _ = priors.sample_array(64, np.random.default_rng(3), cols)
    ref = harness.oracle_logl(olik, fixed, pts, cols)
    eng = lik.sub_model.engine_for(cols)
    launches0 = eng.get_info("launches")
    for rep in range(4):                                            # un-captured, capture, replay, replay -- different rows each
        for i in range(5):
            got = eng.logl_host(pts[5 * rep + i:5 * rep + i + 1])
            assert_logl_close(got, ref[5 * rep + i:5 * rep + i + 1])
        assert_logl_close(eng.logl_host(pts[:17]), ref[:17])        # another size in between
    assert eng.get_info("last_path") == 5
    assert eng.get_info("launches") - launches0 == 2 * (4 * 5 + 4)  # replays count the kernels they launch
    assert lik.log_likelihood(dict(zip(cols, pts[40]))) == pytest.approx(ref[40], rel=1e-4)
    # an option change drops the graphs; the next calls go through another kernel family and then recapture
    eng.set_option("path", 3)
    assert_logl_close(eng.logl_host(pts[:1]), ref[:1])
    eng.set_option("path", 0)
    for _ in range(3):
        assert_logl_close(eng.logl_host(pts[:1]), ref[:1])
    # a new configuration (finite detection limit brighter than some detections: every point fails) must not replay the old graph
    lik2, olik2, fixed2, cols2 = build_pair(core, "Bu2019lm", filters, filters, lc_data, priors, detection_limit=19.0)
    eng2 = lik2.sub_model.engine_for(cols2)
    for _ in range(3):
        assert eng2.logl_host(pts[:1])[0] == SENTINEL
    eng.set_option("cuda_graphs", 0)
    assert_logl_close(eng.logl_host(pts[:3]), ref[:3])
