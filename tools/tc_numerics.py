#!/usr/bin/env python
"""CPU emulation of tensor-core operand roundings for the surrogate MLP (evidence for DESIGN.md section 6).

For the real Bu2019nsbh fixture networks and the random-init Bu2019lm-shaped networks it evaluates the two Dense
layers with (a) fp32 FFMA-like arithmetic, (b) single-pass TF32 operands, (c) single-pass BF16 operands, (d) the
3-pass TF32 split  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (operands truncated to TF32 the way tcgen05 kind::tf32
reads them; the round-1 kernel), (e) the fp16 hi/lo split of the round-2 kernel ("f16x3": kind::f16 operands brought into the
fp16 range by the exact power-of-two scalings of csrc/tc_kernel.cuh / api.cu -- rows of [W1; b1], columns of W2, one scale
per point -- h_hi rounded toward zero with the ReLU, remainders rounded to nearest, subnormals kept), all with fp32 accumulation, and reports the coefficient error against an fp64 evaluation and the
magnitude error it maps to through  |VA[:, :K] . dc| * (maxs - mins).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tf32_trunc(a):
    a = np.ascontiguousarray(a, np.float32)
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16_rn(a):
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split_tf32(a):
    hi = tf32_trunc(a)
    lo = (np.asarray(a, np.float32) - hi).astype(np.float32)
    return hi, tf32_trunc(lo)


def mm32(a, b):
    """fp32 product with fp32 accumulation (operands already rounded)."""
    return (a.astype(np.float32) @ b.astype(np.float32)).astype(np.float32)


def f16_rn(a):
    with np.errstate(over="ignore"):
        return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def f16_rz(a):
    a = np.asarray(a, np.float32)
    with np.errstate(over="ignore"):
        h = a.astype(np.float16)
    over = np.abs(h.astype(np.float32)) > np.abs(a)
    return np.where(over, np.nextafter(h, np.float16(0)), h).astype(np.float32)


def pow2_scale(vmax, top):
    """2^p with vmax 2^p in [2^(top-1), 2^top); 1 for zeros (api.cu: pow2_scale)."""
    vmax = np.asarray(vmax, np.float64)
    safe = np.where(vmax > 0, vmax, 1.0)
    return np.where(vmax > 0, np.exp2(top - (np.floor(np.log2(safe)) + 1)), 1.0).astype(np.float32)


def mlp_f16x3(x, W1, b1, W2, b2, chunk=64, group=4):
    """The round-2 tensor-core front end, operation by operation (products of two fp16 values are exact in fp32; the
    accumulator's round-toward-zero adds are not emulated: NumPy adds round to nearest)."""
    Wa = np.concatenate([W1, b1[None, :]], 0).astype(np.float32)              # rows = inputs, bias
    rs = pow2_scale(np.abs(Wa).max(1), 10)
    Ws = (Wa * rs[:, None]).astype(np.float32)
    W1h = f16_rn(Ws); W1l = f16_rn(Ws - W1h)
    xa = np.concatenate([x.astype(np.float32), np.ones((len(x), 1), np.float32)], 1) / rs[None, :]
    S = np.abs(xa).sum(1)
    sc = np.exp2(3 - np.floor(np.log2(S))).astype(np.float32)                 # 2^e S in [8, 16)
    a = (xa * sc[:, None]).astype(np.float32)
    ah = f16_rn(a); al = f16_rn(a - ah)
    v = (mm32(ah, W1h) + mm32(al, W1h) + mm32(ah, W1l)).astype(np.float32)
    hh = f16_rz(np.maximum(v, 0))
    assert np.isfinite(hh).all() and hh.max() < 2.0 ** 14
    hl = f16_rn(np.maximum(v - hh, 0))
    cs = pow2_scale(np.abs(W2).max(0), 4)
    Ws2 = (W2 * cs[None, :]).astype(np.float32)
    W2h = f16_rn(Ws2); W2l = f16_rn((Ws2 - W2h) * np.float32(2048.0))
    acc_h = np.zeros((len(x), W2.shape[1]), np.float32); acc_x = np.zeros_like(acc_h)
    for j0 in range(0, W2.shape[0], chunk * group):                           # group partials, summed by the CUDA cores
        sl = slice(j0, j0 + chunk * group)
        acc_h += (mm32(hh[:, sl], W2h[sl]) + mm32(hl[:, sl], W2h[sl])).astype(np.float32)
        acc_x += mm32(hh[:, sl], W2l[sl])
    out = (acc_h + acc_x * np.float32(2.0 ** -11)) * (1.0 / sc)[:, None] * (1.0 / cs)[None, :]
    return out.astype(np.float32) + b2


def mlp_variants(x, W1, b1, W2, b2):
    x32 = x.astype(np.float32)
    xa = np.concatenate([x32, np.ones((len(x32), 1), np.float32)], 1)        # bias folded as an extra input
    Wa = np.concatenate([W1, b1[None, :]], 0).astype(np.float32)
    out = {}
    truth = np.maximum(xa.astype(np.float64) @ Wa.astype(np.float64), 0) @ W2.astype(np.float64) + b2
    out["fp32"] = mm32(np.maximum(mm32(xa, Wa), 0), W2) + b2
    h = np.maximum(mm32(tf32_trunc(xa), tf32_trunc(Wa)), 0)
    out["tf32x1"] = mm32(tf32_trunc(h), tf32_trunc(W2)) + b2
    h = np.maximum(mm32(bf16_rn(xa), bf16_rn(Wa)), 0)
    out["bf16x1"] = mm32(bf16_rn(h), bf16_rn(W2)) + b2
    xh, xl = split_tf32(xa); wh, wl = split_tf32(Wa)
    h = np.maximum(mm32(xh, wh) + mm32(xl, wh) + mm32(xh, wl), 0).astype(np.float32)
    hh, hl = split_tf32(h); vh, vl = split_tf32(W2)
    out["tf32x3"] = (mm32(hh, vh) + mm32(hl, vh) + mm32(hh, vl)).astype(np.float32) + b2
    out["f16x3"] = mlp_f16x3(x, W1, b1, W2, b2)
    # layer 1 exact fp32 (CUDA cores), layer 2 single-pass TF32: the cheapest mixed variant
    h = np.maximum(mm32(xa, Wa), 0)
    out["l2_tf32x1"] = mm32(tf32_trunc(h), tf32_trunc(W2)) + b2
    return truth, out


def report(name, nets, x):
    print(f"\n== {name}: {len(x)} points ==")
    print(f"{'variant':10s} {'max|dc|':>10s} {'max dmag':>10s} {'p99 dmag':>10s}")
    agg = {}
    for filt, (W1, b1, W2, b2, VA, rng_) in nets.items():
        truth, out = mlp_variants(x, W1, b1, W2, b2)
        for k, v in out.items():
            dc = v.astype(np.float64) - truth
            dmag = np.abs(dc @ VA.T) * rng_[None, :]
            a = agg.setdefault(k, [0.0, 0.0, []])
            a[0] = max(a[0], np.abs(dc).max()); a[1] = max(a[1], dmag.max()); a[2].append(np.percentile(dmag, 99))
    for k, (dc, dm, p99) in agg.items():
        print(f"{k:10s} {dc:10.2e} {dm:10.2e} {max(p99):10.2e}")
    return agg


def main():
    rng = np.random.default_rng(7)
    z = np.load(os.path.join(ROOT, "tests", "golden", "bu2019nsbh_fixture.npz"), allow_pickle=True)
    nets = {}
    for f in ("ztfr", "sdssu", "2massks"):
        K = z[f + "/W2"].shape[1]
        nets[f] = (z[f + "/W1"], z[f + "/b1"], z[f + "/W2"], z[f + "/b2"], z[f + "/VA"][:, :K],
                   z[f + "/maxs"] - z[f + "/mins"])
    x = rng.uniform(-0.1, 1.5, size=(4096, 3))          # scaled inputs incl. the extrapolation the golden test uses
    report("Bu2019nsbh fixture (trained Keras weights)", nets, x)
    from nmma_b200 import synthetic as syn
    core = syn.random_model("Bu2019lm", syn.AT2017GFO_FILTERS, seed=0)
    nets = {}
    for f, e in core.items():
        W1, b1, W2, b2 = e["model"]
        nets[f] = (W1, b1, W2, b2, e["VA"][:, :10], e["maxs"] - e["mins"])
    x = rng.uniform(-0.6, 1.2, size=(4096, 4))
    report("Bu2019lm-shaped random init (bench workload)", nets, x)


if __name__ == "__main__":
    main()
