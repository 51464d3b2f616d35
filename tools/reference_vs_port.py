#!/usr/bin/env python
"""How fair is the CPU arm?  Times the reference's OWN classes (nmma.em.model.SVDLightCurveModel +
nmma.em.em_likelihood.EMTransientLikelihood from /root/reference, third-party imports stubbed as in
tests/golden/reference_stubs.py) against the oracle PORT that `bench.py --impl reference` runs, on the same points, one
core, in this container.  (/root/reference does not exist on the GPU box, so the bench arm itself runs the port.)

    python tools/reference_vs_port.py [n_points]
"""
import copy, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
warnings.filterwarnings("ignore")
import numpy as np
import reference_stubs as RS
import make_reference_vectors as MV
from helpers import build_reference_pair
from oracle import harness

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
mods = RS.load_reference()
from nmma_b200.core import priors as P
filters = ["ztfr", "sdssu", "2massks"]
model = mods["model"].SVDLightCurveModel("Bu2019nsbh", svd_path=MV.DATA, interpolation_type="tensorflow", filters=list(filters),
                                         local_only=True)
raw = MV.observations(filters, seed=ord("A"))
lc_data = mods["utils"].setup_filtered_lc_data(copy.deepcopy(raw), 57000.0)
priors = MV.base_priors(P)
handler = mods["systematics"].FilterSystematicsHandler(list(filters), None, 1.0, lc_data[0])
lik = mods["em_likelihood"].EMTransientLikelihood(model, lc_data, handler, priors, filters=list(filters), detection_limit=np.inf)
cols = list(priors.keys())
pts, _ = priors.sample_array(n, np.random.default_rng(0), cols)
for row in pts[:20]:
    lik.log_likelihood(dict(zip(cols, map(float, row))))
t0 = time.perf_counter()
ref = np.array([lik.log_likelihood(dict(zip(cols, map(float, row)))) for row in pts])
t_ref = time.perf_counter() - t0
_, olik, fixed, ocols, _ = build_reference_pair("A")
harness.oracle_logl(olik, fixed, pts[:20], ocols)
t0 = time.perf_counter()
port = harness.oracle_logl(olik, fixed, pts, ocols)
t_port = time.perf_counter() - t0
err = np.abs(port - ref) / np.maximum(1, np.abs(ref))
print(f"reference classes (stubbed Keras/astropy): {n / t_ref:8.1f} evals/s one core ({1e3 * t_ref / n:.3f} ms/eval)")
print(f"oracle port (bench.py --impl reference):   {n / t_port:8.1f} evals/s one core ({1e3 * t_port / n:.3f} ms/eval)")
print(f"port / reference speed = {t_ref / t_port:.2f}x; max rel |dlogL| = {err.max():.2e} (3 filters, Bu2019nsbh fixture, case A)")
