#!/usr/bin/env python
"""GPU check of the hybrid kernel (path 4) against the FFMA fused kernel (1), the two-stage kernels (2), the TC kernel
(3) and the oracle on the bench workload; then a quick timing.  Run under gpurun with a timeout."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
wl = bench.build_workload()
cols = list(wl["priors"].keys())
lik, model, handler = bench.gpu_likelihood(wl, 0)
eng = lik.sub_model.engine_for(cols)
print("hy_supported", eng.get_info("hy_supported"), flush=True)
pts, _ = wl["priors"].sample_array(N, np.random.default_rng(5), cols)
dev = torch.from_numpy(pts).cuda()
res = {}
for path in (2, 1, 3, 4):
    eng.set_option("path", path)
    out = eng.logl_device(dev)
    torch.cuda.synchronize()
    res[path] = out.cpu().numpy()
    print("path", path, "done", res[path][:3], flush=True)
def rel(a, b):
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))
for p in (1, 3, 4):
    print(f"max rel {p} vs 2:", rel(res[p], res[2]).max(), "argmax", rel(res[p], res[2]).argmax())
print("sentinel rows equal:", np.array_equal(res[4] < -1e300, res[2] < -1e300), int((res[2] < -1e300).sum()))
from oracle import harness
olik, fixed = harness.build_oracle_likelihood(wl["core"], model.model_parameters, wl["filters"],
                                              np.asarray(model.model_times, float), wl["filters"], wl["lc_data"],
                                              wl["priors"], sys_plan=handler.device_plan(), z_table=model._z_table)
ref = harness.oracle_logl(olik, fixed, pts[:200], cols)
for p in (1, 3, 4):
    print(f"max rel {p} vs oracle (200):", rel(res[p][:200], ref).max())
M = 1_000_000
big, _ = wl["priors"].sample_array(M, np.random.default_rng(6), cols)
bigd = torch.from_numpy(big).cuda()
out = torch.empty(M, dtype=torch.float64, device="cuda")
for path in (1, 3, 4):
    eng.set_option("path", path)
    for _ in range(2):
        eng.logl_device(bigd, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.logl_device(bigd, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"path {path}: {ms:.3f} ms per 1e6 evals = {M / ms / 1e3:.1f} M evals/s", flush=True)
    res[("big", path)] = out.cpu().numpy().copy()
print("1e6 points: max rel 4 vs 1:", rel(res[("big", 4)], res[("big", 1)]).max())
print("1e6 points: max rel 3 vs 1:", rel(res[("big", 3)], res[("big", 1)]).max())
