#!/usr/bin/env python
"""Device time of the two-stage GP kernels (path 2) and the fused GP kernel (path 4) by batch size (config 4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
wl = bench.make_spec("c4"); cols = wl["cols"]
lik = bench.gpu_likelihood(wl, 0)
eng = lik.sub_model.engine_for(cols)
for M in (1, 32, 128, 512, 1024, 2048, 4096, 8192, 16384, 65536, 200000):
    big, _ = wl["priors"].sample_array(M, np.random.default_rng(6), cols)
    bigd = torch.from_numpy(big).cuda(); out = torch.empty(M, dtype=torch.float64, device="cuda")
    line = f"N = {M:6d}:"
    for path in (2, 4):
        eng.set_option("path", path)
        for _ in range(2): eng.logl_device(bigd, out=out)
        torch.cuda.synchronize()
        reps = 20 if M <= 16384 else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): eng.logl_device(bigd, out=out)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        line += f"  path {path} {us:9.1f} us ({M / us:7.2f} M evals/s)"
    print(line, flush=True)
