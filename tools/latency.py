#!/usr/bin/env python
"""Latency of the one-point-per-call seam (bilby -> pymultinest calls log_likelihood(dict) sequentially,
nmma/core/base.py:77-82) and of small batches through the host-buffer entry: python tools/latency.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
wl = bench.make_spec("c2"); cols = wl["cols"]
lik = bench.gpu_likelihood(wl, 0)
pts, _ = wl["priors"].sample_array(4096, np.random.default_rng(3), cols)
dicts = [dict(zip(cols, row)) for row in pts[:2000]]
for d in dicts[:50]: lik.log_likelihood(d)
t = time.perf_counter()
for d in dicts: lik.log_likelihood(d)
dt = (time.perf_counter() - t) / len(dicts)
print(f"log_likelihood(dict), one point per call: {dt * 1e6:.1f} us per call = {1 / dt:.0f} evals/s", flush=True)
for n in (1, 16, 256, 1024, 4096):
    x = np.ascontiguousarray(pts[:n])
    for _ in range(20): lik.log_likelihood_batch(x, cols)
    t = time.perf_counter()
    for _ in range(200): lik.log_likelihood_batch(x, cols)
    dt = (time.perf_counter() - t) / 200
    print(f"log_likelihood_batch, N = {n:5d} (host buffers): {dt * 1e6:.1f} us per call = {n / dt:.3g} evals/s", flush=True)
# which kernel family is fastest at which batch size (device-resident points, CUDA events)
eng = lik.sub_model.engine_for(cols)
big, _ = wl["priors"].sample_array(262144, np.random.default_rng(4), cols)
bigd = torch.from_numpy(big).cuda()
print("device-resident, us per call by path (1 = fused FFMA, 2 = two-stage, 3 = tensor core, 5 = hidden-split tensor core + back end):")
for n in (1, 64, 256, 512, 1024, 2048, 4096, 16384, 65536, 262144):
    row = []
    for path in (1, 2, 3, 5):
        eng.set_option("path", path)
        x = bigd[:n].contiguous(); out = torch.empty(n, dtype=torch.float64, device="cuda")
        try:
            for _ in range(5): eng.logl_device(x, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): eng.logl_device(x, out=out)
            e1.record(); torch.cuda.synchronize()
            row.append(f"{e0.elapsed_time(e1) / 20 * 1e3:9.1f}")
        except Exception as ex:
            row.append("      n/a")
    print(f"  N = {n:6d}: " + " ".join(row), flush=True)
eng.set_option("path", 0)
