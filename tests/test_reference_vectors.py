"""The oracle (and on the GPU box the CUDA path) against outputs of the reference's own source files.

tests/golden/reference_vectors.npz was produced by running nmma.em.em_likelihood.EMTransientLikelihood
etc. unmodified from /root/reference (tests/golden/make_reference_vectors.py).  This pins the
log-likelihood part of the path, which no test of the reference itself asserts.
"""
import numpy as np
import pytest

from helpers import REFERENCE_CASES, SENTINEL, assert_logl_close, build_reference_pair


@pytest.mark.parametrize("name", list(REFERENCE_CASES))
def test_oracle_reproduces_reference_logl(name):
    from oracle import harness
    lik, olik, fixed, cols, gold = build_reference_pair(name)
    n = 96 if name != "D" else 32
    got = harness.oracle_logl(olik, fixed, gold["points"][:n], cols)
    ref = gold["logl"][:n]
    assert (ref == SENTINEL).sum() >= (0 if n < 96 else 5)
    err = assert_logl_close(got, ref, rtol=1e-10)        # same NumPy / SciPy / sklearn calls: round-off only
    print(name, "oracle vs reference max rel err", err)


@pytest.mark.parametrize("name", ["A", "B", "D"])
def test_oracle_reproduces_reference_mags(name):
    lik, olik, fixed, cols, gold = build_reference_pair(name)
    model = olik.light_curve_model
    for i in range(4):
        p = dict(fixed)
        p.update(dict(zip(cols, gold["points"][i])))
        p = model.parameter_conversion(p)
        tobs, lc = model.gen_detector_lc(p)
        assert np.array_equal(tobs, gold["tobs"][i])
        for fi, f in enumerate(model.filters):
            a, b = lc[f], gold["mags"][i][fi]
            assert np.array_equal(np.isfinite(a), np.isfinite(b))
            assert np.allclose(a[np.isfinite(a)], b[np.isfinite(b)], rtol=0, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(REFERENCE_CASES))
def test_cuda_reproduces_reference_logl(name):
    import torch
    assert torch.cuda.is_available()
    lik, olik, fixed, cols, gold = build_reference_pair(name)
    eng = lik.sub_model.engine_for(cols)
    ref = gold["logl"]
    paths = {"auto": 0, "two_stage": 2}
    if eng.get_info("fused_supported"):
        paths["fused"] = 1
    for pname, path in paths.items():
        eng.set_option("path", path)
        got = lik.log_likelihood_batch(gold["points"], cols)
        print(name, pname, "CUDA vs reference max rel err", assert_logl_close(got, ref, rtol=1e-4))
    eng.set_option("path", 0)
    one = lik.log_likelihood(dict(zip(cols, gold["points"][3])))
    assert one == pytest.approx(ref[3], rel=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["A", "B", "D"])
def test_cuda_reproduces_reference_mags(name):
    import torch
    assert torch.cuda.is_available()
    lik, olik, fixed, cols, gold = build_reference_pair(name)
    model = lik.sub_model.light_curve_model
    worst = 0.0
    for i in range(8):
        p = dict(fixed)
        p.update(dict(zip(cols, gold["points"][i])))
        p = model.parameter_conversion(p)
        tobs, lc = model.gen_detector_lc(p)
        assert np.array_equal(tobs, gold["tobs"][i]), "detector-frame grid must be bit-exact"
        for fi, f in enumerate(model.filters):
            a, b = lc[f], gold["mags"][i][fi]
            assert np.array_equal(np.isfinite(a), np.isfinite(b)), "finite mask must be bit-exact"
            worst = max(worst, np.abs(a[np.isfinite(a)] - b[np.isfinite(b)]).max())
    assert worst < 1e-3
    print(name, "CUDA vs reference max |dmag|", worst)
