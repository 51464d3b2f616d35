#!/usr/bin/env python
"""Time one path on the bench workload: python tools/tc_time.py PATH [M] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
path = int(sys.argv[1]); M = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = os.environ.get("TC_TIME_CONFIG", "c2")
wl = bench.make_spec(cfg); cols = wl["cols"]
lik = bench.gpu_likelihood(wl, 0)
eng = lik.sub_model.engine_for(cols); eng.set_option("path", path)
big, _ = wl["priors"].sample_array(M, np.random.default_rng(6), cols)
bigd = torch.from_numpy(big).cuda(); out = torch.empty(M, dtype=torch.float64, device="cuda")
for _ in range(2): eng.logl_device(bigd, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): eng.logl_device(bigd, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"path {path}: {ms:.3f} ms per {M} evals = {M / ms / 1e3:.1f} M evals/s", flush=True)
