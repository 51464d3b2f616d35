#!/usr/bin/env python
"""Where the time of a one-point call goes (bilby -> pymultinest seam): python tools/latency_breakdown.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
wl = bench.make_spec("c2"); cols = wl["cols"]
lik = bench.gpu_likelihood(wl, 0)
eng = lik.sub_model.engine_for(cols)
pts, _ = wl["priors"].sample_array(64, np.random.default_rng(3), cols)
row = np.ascontiguousarray(pts[:1]); out = np.empty(1)
d = dict(zip(cols, pts[0]))


def t(f, n=3000):
    for _ in range(100): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6


print("lik.log_likelihood(dict)          %.1f us" % t(lambda: lik.log_likelihood(d)))
print("sub_model.log_likelihood(dict)    %.1f us" % t(lambda: lik.sub_model.log_likelihood(d)))
print("eng.logl_host(row, out)           %.1f us" % t(lambda: eng.logl_host(row, out=out)))
pin = torch.from_numpy(row.copy()).pin_memory(); pout = torch.empty(1, dtype=torch.float64).pin_memory()
pn, po = pin.numpy(), pout.numpy()
print("eng.logl_host(pinned row, pinned) %.1f us" % t(lambda: eng.logl_host(pn, out=po)))
dev = torch.from_numpy(row).cuda(); dout = torch.empty(1, dtype=torch.float64, device="cuda")


def devcall():
    eng.logl_device(dev, out=dout); torch.cuda.synchronize()


print("eng.logl_device + synchronize     %.1f us" % t(devcall))
for path in (2, 3):
    eng.set_option("path", path)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(20): eng.logl_device(dev, out=dout)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200): eng.logl_device(dev, out=dout)
    e1.record(); torch.cuda.synchronize()
    print("path %d, N = 1, back-to-back device time %.1f us" % (path, e0.elapsed_time(e1) / 200 * 1e3))
