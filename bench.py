#!/usr/bin/env python
"""Benchmark: kilonova logL evals/sec (Bu2019lm vs AT2017gfo), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--points M]

A "step" is one pass of the hot path (fused surrogate + likelihood) over one batch of M = 10^6
prior draws from priors/Bu2019lm.prior per GPU (configs[1] of BASELINE.json; under torchrun every
rank evaluates its own 10^6-point shard and the logL blocks are all-gathered over NCCL: weak
scaling, configs[4]).  Weights are random-init of the Bu2019lm architecture (the Zenodo weights
are not available offline); photometry is the real AT2017gfo table cut at 14 d.

JSON line keys (base contract + tier additions): value (device-resident inputs), e2e (host buffers
through EMTransientLikelihood.log_likelihood_batch, H2D/D2H inside the timed region), roofline
(FP32-FMA compute bound; measured FFMA peak as denominator, HBM figures alongside), cpu_baseline
(the oracle port on the box's host cores, bounded sample), clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "kilonova logL evals/sec (Bu2019lm vs AT2017gfo)"
UNIT = "evals/s"
WORKLOAD = "Bu2019lm batched likelihood sweep: 10^6 prior draws from priors/Bu2019lm.prior vs AT2017gfo (data_tmax 14 d, 133 obs, 9 filters)"
N_ROTATE = 6  # distinct input batches cycled between steps: 6 x 48 MB = 288 MB > 126 MB L2


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def build_workload():
    from nmma_b200 import synthetic as syn
    lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
    core = syn.random_model("Bu2019lm", filters, seed=0)
    priors = syn.bu2019lm_prior()
    return dict(lc_data=lc_data, filters=filters, core=core, priors=priors)


def gpu_likelihood(wl, device):
    from nmma_b200.em import EMTransientLikelihood, FilterSystematicsHandler, SVDLightCurveModel
    model = SVDLightCurveModel("Bu2019lm", svd_mag_model=wl["core"], interpolation_type="tensorflow",
                               filters=wl["filters"], device=device)
    handler = FilterSystematicsHandler(wl["filters"], None, 1.0, wl["lc_data"][0])
    lik = EMTransientLikelihood(model, wl["lc_data"], handler, wl["priors"], filters=wl["filters"])
    return lik, model, handler


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference's per-point Python path) on the host cores
# ---------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(wl_blob):
    try:
        from threadpoolctl import threadpool_limits
        _W["tl"] = threadpool_limits(1)          # NMMA pins BLAS threads to 1 for pooled runs (joint/main.py:2)
    except Exception:  # noqa: BLE001
        pass
    import warnings
    warnings.filterwarnings("ignore")
    from nmma_b200.em.model import model_parameters_dict
    from oracle import harness
    wl = wl_blob
    tt = next(iter(wl["core"].values()))["tt"]
    lik, fixed = harness.build_oracle_likelihood(
        wl["core"], model_parameters_dict["Bu2019lm"], wl["filters"], np.asarray(tt, float), wl["filters"],
        wl["lc_data"], wl["priors"], sys_plan=None, error_budget=1.0, z_table=wl["z_table"])
    _W["lik"], _W["fixed"], _W["cols"] = lik, fixed, wl["cols"]


def _cpu_eval(chunk):
    from oracle import harness
    return harness.oracle_logl(_W["lik"], _W["fixed"], chunk, _W["cols"])


class CpuArm:
    """multiprocessing pool over the host cores (the analogue of the reference's schwimmbad task farm,
    nmma/core/mpi_setup.py:651-683); workers are forked before CUDA is initialised."""

    def __init__(self, wl, cols, z_table, cores=None):
        import multiprocessing as mp
        self.cores = cores or len(os.sched_getaffinity(0))
        blob = dict(wl)
        blob["cols"], blob["z_table"] = cols, z_table
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init, initargs=(blob,))

    def run(self, pts):
        chunks = np.array_split(pts, self.cores * 4)
        t0 = time.perf_counter()
        out = np.concatenate(self.pool.map(_cpu_eval, chunks))
        return out, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def z_table_for(priors):
    from nmma_b200.core.conversion import get_cosmo_grids
    dl = priors["luminosity_distance"]
    return get_cosmo_grids(dl.minimum, dl.maximum)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7 or not (t0 <= t <= t1 + 0.1):
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle port; the real package cannot be
    imported offline -- DESIGN.md) on all host cores; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = build_workload()
    cols = [k for k in wl["priors"].keys()]
    arm = CpuArm(wl, cols, z_table_for(wl["priors"]))
    rng = np.random.default_rng(1234)
    probe, _ = wl["priors"].sample_array(arm.cores * 40, rng, cols)
    _, dt = arm.run(probe)
    rate = len(probe) / dt
    # bounded sample: ~4 s of CPU work per step, shrunk so that steps + warm-up stay within ~2 minutes whatever K is
    sec_per_step = min(4.0, max(0.2, 120.0 / max(1, args.steps + args.warmup)))
    per_step = int(max(arm.cores * 40, min(rate * sec_per_step, 200000)))
    pts, _ = wl["priors"].sample_array(per_step, rng, cols)
    for _ in range(args.warmup):
        arm.run(pts)
    t = 0.0
    for _ in range(args.steps):
        _, dt = arm.run(pts)
        t += dt
    arm.close()
    value = per_step * args.steps / t
    sample = f"{per_step} prior draws per step through the per-point Python path (NumPy/SciPy oracle port, fp32 NumPy MLP standing in for Keras)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 MLP / f64 likelihood", "data": "synthetic weights, real AT2017gfo photometry",
        "config": {"workload": WORKLOAD, "points_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    M = args.points
    wl = build_workload()
    cols = [k for k in wl["priors"].keys()]
    ztab = z_table_for(wl["priors"])

    # ---- CPU baseline first (fork before CUDA init), rank 0 at N = 1 only ----
    cpu = None
    if world == 1 and not args.no_cpu:
        arm = CpuArm(wl, cols, ztab)
        rng = np.random.default_rng(99)
        probe, _ = wl["priors"].sample_array(arm.cores * 40, rng, cols)
        _, dt = arm.run(probe)
        n_s = int(max(arm.cores * 40, min(len(probe) / dt * 12.0, 400000)))      # ~12 s of CPU work
        cpu_pts, _ = wl["priors"].sample_array(n_s, rng, cols)
        cpu_ref, dt = arm.run(cpu_pts)
        one_core_pts = cpu_pts[:300]
        _cpu_init(dict(wl, cols=cols, z_table=ztab))
        t0 = time.perf_counter()
        _cpu_eval(one_core_pts)
        one_core = len(one_core_pts) / (time.perf_counter() - t0)
        arm.close()
        cpu = {"value": n_s / dt, "unit": UNIT, "cores": arm.cores, "kind": "port",
               "sample": f"{n_s} of the same prior draws through the per-point NumPy/SciPy oracle port "
                         f"(fp32 NumPy MLP standing in for Keras), multiprocessing pool on all host cores",
               "one_core_value": one_core, "cpu_count": os.cpu_count()}

    # nvidia-smi needs ~1 s to start: launch it now, keep the samples that fall inside the timed regions
    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    lik, model, handler = gpu_likelihood(wl, local_rank)
    eng = lik.sub_model.engine_for(cols)
    if args.path:
        eng.set_option("path", args.path)
    if args.packed is not None:
        eng.set_option("packed_fma", args.packed)
    if args.pt is not None:
        eng.set_option("points_per_thread", args.pt)

    rng = np.random.default_rng(1234 + rank)
    host_batches = []
    for _ in range(N_ROTATE):
        p, _ = wl["priors"].sample_array(M, rng, cols)
        host_batches.append(torch.from_numpy(p).pin_memory())
    dev_batches = [b.to(dev) for b in host_batches]
    out_local = torch.empty(M, dtype=torch.float64, device=dev)
    out_all = torch.empty(M * world, dtype=torch.float64, device=dev) if world > 1 else out_local
    out_host = torch.empty(M * world if rank == 0 else 1, dtype=torch.float64).pin_memory()

    # parity spot check against the CPU oracle inside the same run
    parity = None
    if cpu is not None:
        got = lik.log_likelihood_batch(cpu_pts[:2000], cols)
        err = np.abs(got - cpu_ref[:2000]) / np.maximum(1.0, np.abs(cpu_ref[:2000]))
        parity = {"points": 2000, "max_rel_err": float(err.max()), "tolerance": 1e-4}
        assert err.max() < 1e-4, f"GPU/CPU logL disagree: {err.max()}"

    ffma = None
    if rank == 0:
        ffma = {"scalar_tflops": eng.ffma_peak(0, 20000) / 1e12, "packed_tflops": eng.ffma_peak(1, 20000) / 1e12}

    def step_device(i):
        eng.logl_device(dev_batches[i % N_ROTATE], out=out_local)
        if world > 1:
            dist.all_gather_into_tensor(out_all, out_local)

    def step_e2e(i):
        hb = host_batches[i % N_ROTATE]
        if world == 1:
            lik.log_likelihood_batch(hb.numpy(), cols, out=out_host.numpy())
        else:
            eng.logl_host(hb.numpy(), out=out_local)       # pipelined H2D + kernels, result stays on the device
            dist.all_gather_into_tensor(out_all, out_local)
            if rank == 0:
                out_host.copy_(out_all, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, per_launch=False):
        for i in range(warmup):
            fn(i)
        sync()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            if per_launch:
                evs[i][0].record()
            fn(warmup + i)
            if per_launch:
                evs[i][1].record()
        e1.record()
        sync()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per = [a.elapsed_time(b) for a, b in evs] if per_launch else None
        return float(t.item()), per, (t0, t1)

    l0 = eng.get_info("launches")
    ms_total, per_launch, (w0, w1) = timed(step_device, args.steps, args.warmup, per_launch=True)
    launches = eng.get_info("launches") - l0 - args.warmup

    # e2e: wall-clock inside the C call includes the copies; device events would miss the host part
    for i in range(max(args.warmup, 3)):
        step_e2e(i)
    sync()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    sync()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    w2 = time.perf_counter()
    clocks = None
    if sampler:
        clocks = sampler.stop(w0, w2)          # device-timed loop + e2e loop, both under load
        if clocks is not None:
            clocks["window"] = "device-timed steps and e2e steps (contiguous, GPU busy throughout)"

    if rank == 0:
        total_pts = M * world
        value = total_pts * args.steps / (ms_total * 1e-3)
        flop = eng.get_info("algorithmic_flop_per_eval")
        last_path = eng.get_info("last_path")
        kern_ms = float(np.mean(per_launch))
        achieved = M * flop / (kern_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        P = len(cols)
        ffma_peak = max(ffma["scalar_tflops"], ffma["packed_tflops"])
        hbm = {"achieved_gbs": M * (P * 8 + 8) / (kern_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
               "frac": M * (P * 8 + 8) / (kern_ms * 1e-3) / 1e9 / hbm_peak,
               "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"}
        if last_path == 3:
            # tensor-core kernel: tcgen05 kind::tf32 runs at half the bf16 rate (tools/tc_probe.cu: 128*N/256 cycles per
            # K=8 MMA = 4096 FLOP/clk/SM); the 3xTF32 split and the N=16 / K=8 operand padding make the EXECUTED
            # tensor FLOPs 4.79x the algorithmic ones -- both are reported, `achieved` stays algorithmic.
            bf16_peak = peaks.get("bf16_tflops", 1650.0)
            tf32_peak = bf16_peak / 2.0
            executed = eng.get_info("tc_executed_flop_per_eval")
            roofline = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                        "frac": achieved / tf32_peak,
                        "peak_source": ("MEASURED_PEAKS.json bf16_tflops / 2 (dense tf32 = half the bf16 rate)" if peaks
                                        else "fallback 1650 bf16 TFLOP/s / 2"),
                        "executed_tensor_tflops": M * executed / (kern_ms * 1e-3) / 1e12,
                        "executed_frac_of_tf32_peak": M * executed / (kern_ms * 1e-3) / 1e12 / tf32_peak,
                        "executed_tensor_flop_per_eval": executed,
                        "frac_of_ffma_peak": achieved / ffma_peak}
        else:
            roofline = {"bound": "fp32_fma (CUDA cores; neither HBM nor tensor: SURVEY.md 8d)",
                        "achieved": achieved, "peak": ffma_peak, "unit": "TFLOP/s", "frac": achieved / ffma_peak,
                        "peak_source": "FFMA micro-benchmark measured in this run (nmma_b200_ffma_peak); nominal 74.4 TFLOP/s at 1965 MHz",
                        "frac_of_nominal": achieved / 74.4}
        roofline.update({"ffma_peak_measured": ffma, "algorithmic_flop_per_eval": flop, "kernel_ms_per_launch": kern_ms,
                         "hbm": hbm, "traffic": traffic})
        kname = {1: "fused_mlp_logl_kernel (FFMA)", 2: "two_stage", 3: "fused_tc_logl_kernel (tcgen05 3xTF32)"}.get(last_path, "?")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 MLP (3xTF32 split on tcgen05) / f64 likelihood" if last_path == 3 else "f32 MLP (FFMA) / f64 likelihood",
            "data": "synthetic: random-init Bu2019lm-shaped weights, real AT2017gfo photometry",
            "config": {"workload": WORKLOAD, "points_per_gpu_per_step": M, "global_points_per_step": total_pts,
                       "sharding": f"contiguous row blocks x{world}, NCCL all-gather of logL" if world > 1 else "single GPU",
                       "l2": f"inputs rotate through {N_ROTATE} distinct batches ({N_ROTATE * M * P * 8 / 1e6:.0f} MB > 126 MB L2)",
                       "kernel_path": kname},
            "e2e": {"value": total_pts * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": total_pts * P * 8, "d2h_bytes_per_step": total_pts * 8,
                    "api": "EMTransientLikelihood.log_likelihood_batch -> nmma_b200_logl_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity_in_run": parity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--packed", type=int, default=None)
    ap.add_argument("--pt", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        import __graft_entry__ as ge
        if not os.path.isfile(ge.LIB):
            ge.build()
        gpu_arm(args)


if __name__ == "__main__":
    main()
