#!/usr/bin/env python
"""Benchmark: kilonova logL evals/sec (Bu2019lm vs AT2017gfo), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--points M] [--configs c3,c4,c5]

Headline line (``value`` / ``e2e`` / ``roofline`` / ``cpu_baseline``): BASELINE.json configs[1] -- one "step" is one pass of
the hot path (fused surrogate + likelihood) over M = 10^6 draws from priors/Bu2019lm.prior per GPU against the real
AT2017gfo photometry (cut at 14 d), random-init weights of the Bu2019lm architecture (the Zenodo weights are not available
offline).  Under torchrun every rank evaluates its own 10^6-point block (weak scaling) through
``nmma_b200.sharding.ShardedEvaluator``; the only collective is the NCCL all-gather of the logL blocks, issued on a side
stream under the next step's kernels.  ``e2e`` goes through ``EMTransientLikelihood.log_likelihood_batch`` with HOST buffers
(H2D + kernels + D2H inside the timed region); at N > 1 every rank copies its block into one page-locked shared-memory
result vector (``HostResultBuffer``) that rank 0 reads in place.

``config_lines`` carries short runs of the other BASELINE.json configurations in the same JSON line, each with its rate, a
``roofline`` and an in-run parity check against the CPU oracle:
  c3  configs[2]  Bu2023Ye-shaped (d = 7) + 4 time-node systematics + the 3 AT2017gfo upper limits + detection limit 24.5
  c4  configs[3]  Ka2017-shaped sklearn_gp surrogate (Ntr = 329) across ZTF g/r/i + sdssu + PS1 grizy with detection limits
  c5  configs[4]  10^8 Bu2019lm prior draws made ON the device (Philox, first_index per rank), sharded over the ranks,
                  one NCCL gather of logL[10^8]

``parity_in_run`` compares rows of the TIMED output (device arm: the last timed step's result, every rank's block of the
gathered vector; e2e arm: the shared host vector) with the oracle, at every N.
"""
from __future__ import annotations

import argparse
import copy
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
N_ROTATE = 6      # distinct input batches cycled between steps: 6 x 48 MB = 288 MB > 126 MB L2
N_PARITY = 256    # rows per rank and configuration compared with the oracle
SWEEP_TOTAL = 100_000_000
SWEEP_SEED = 20261018

YAML_TIME = {"config": {"withTime": {"value": True, "filters": [None], "time_nodes": 4, "type": "Uniform", "minimum": 0,
                                     "maximum": 2},
                        "withoutTime": {"value": False, "type": "Uniform", "minimum": 0, "maximum": 2}}}


# ---------------------------------------------------------------------------------------------
# workloads: one picklable spec per configuration, from which the GPU likelihood and the oracle are both built
# ---------------------------------------------------------------------------------------------
def make_spec(name):
    from nmma_b200 import synthetic as syn
    from nmma_b200.em import FilterSystematicsHandler
    if name in ("c2", "c5"):
        lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
        spec = dict(model="Bu2019lm", kind="mlp", filters=filters, core=syn.random_model("Bu2019lm", filters, seed=0),
                    lc_data=lc_data, priors=syn.bu2019lm_prior(), systematics=None, limit=np.inf,
                    workload=WORKLOAD if name == "c2" else
                    f"Bu2019lm {SWEEP_TOTAL:.0e}-point prior-draw sweep, draws made on the device (Philox4x32-10), sharded over the ranks")
    elif name == "c3":
        lc_data, filters = syn.load_at2017gfo(data_tmax=14.0)
        priors = syn.bu2023ye_prior()
        priors["timeshift"].maximum = 0.1
        spec = dict(model="Bu2023Ye", kind="mlp", filters=filters, core=syn.random_model("Bu2023Ye", filters, seed=1),
                    lc_data=lc_data, priors=priors, systematics=copy.deepcopy(YAML_TIME), limit=24.5,
                    workload="Bu2023Ye-shaped (d = 7) + 4 time-node systematics (legacy YAML withTime) + 3 upper limits + "
                             "detection limit 24.5 mag vs AT2017gfo")
    elif name == "c4":
        filters = ["ztfg", "ztfr", "ztfi", "sdssu", "ps1::g", "ps1::r", "ps1::i", "ps1::z", "ps1::y"]
        rng = np.random.default_rng(8)
        times, mags, errs = {}, {}, {}
        for f in filters:
            t = np.sort(rng.uniform(0.3, 13.0, 10))
            e = rng.uniform(0.02, 0.3, 10)
            e[rng.choice(10, size=2, replace=False)] = np.inf
            times[f], mags[f], errs[f] = t, 18.0 + 0.15 * t + rng.normal(scale=0.3, size=10), e
        limits = {"ztfg": 21.7, "ztfr": 21.4, "ztfi": 20.9, "sdssu": 23.9, "ps1::g": 25.0, "ps1::r": 24.7, "ps1::i": 24.0,
                  "ps1::z": 23.3, "ps1::y": 22.1}
        priors = syn.ka2017_prior()
        priors["timeshift"].maximum = 0.2
        spec = dict(model="Ka2017", kind="gp", filters=filters,
                    core=syn.random_model("Ka2017", filters, kind="gp", seed=2, Ntr=329),
                    lc_data=(times, mags, errs, 0.0), priors=priors, systematics=None, limit=limits,
                    workload="Ka2017-shaped sklearn_gp surrogate (Ntr = 329, K = 10) across ZTF g/r/i + sdssu + PS1 grizy, "
                             "10 epochs per filter (2 upper limits each), survey detection limits")
    else:
        raise ValueError(name)
    handler = FilterSystematicsHandler(list(spec["filters"]), spec["systematics"], 1.0, spec["lc_data"][0])
    if spec["systematics"] is not None:
        handler.setup_systematics_priors(spec["priors"])
    spec["name"] = name
    handler.reset(np.asarray(next(iter(spec["core"].values()))["tt"], float), spec["priors"])   # legacy YAML: node times
    spec["sys_plan"] = handler.device_plan()
    spec["cols"] = [k for k, p in spec["priors"].items() if not isinstance(p, (int, float)) and not hasattr(p, "peak")]
    dl = spec["priors"]["luminosity_distance"]
    from nmma_b200.core.conversion import get_cosmo_grids
    spec["z_table"] = get_cosmo_grids(dl.minimum, dl.maximum)
    return spec


def gpu_likelihood(spec, device):
    from nmma_b200.em import EMTransientLikelihood, FilterSystematicsHandler, SVDLightCurveModel
    itype = "sklearn_gp" if spec["kind"] == "gp" else "tensorflow"
    model = SVDLightCurveModel(spec["model"], svd_mag_model=spec["core"], interpolation_type=itype,
                               filters=list(spec["filters"]), device=device)
    handler = FilterSystematicsHandler(list(spec["filters"]), spec["systematics"], 1.0, spec["lc_data"][0])
    lik = EMTransientLikelihood(model, spec["lc_data"], handler, spec["priors"], filters=list(spec["filters"]),
                                detection_limit=spec["limit"])
    assert lik.columns == spec["cols"], (lik.columns, spec["cols"])
    return lik


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference's per-point Python path) on the host cores
# ---------------------------------------------------------------------------------------------
_SPECS = {}      # inherited by the forked workers
_ORACLES = {}


def _cpu_init():
    try:
        from threadpoolctl import threadpool_limits
        _ORACLES["_tl"] = threadpool_limits(1)          # NMMA pins BLAS threads to 1 for pooled runs (joint/main.py:2)
    except Exception:  # noqa: BLE001
        pass
    import warnings
    warnings.filterwarnings("ignore")


def _oracle_for(name):
    if name not in _ORACLES:
        from nmma_b200.em.model import model_parameters_dict
        from oracle import harness
        s = _SPECS[name]
        tt = next(iter(s["core"].values()))["tt"]
        _ORACLES[name] = harness.build_oracle_likelihood(
            harness.oracle_ready_core(s["core"]), model_parameters_dict[s["model"]], s["filters"], np.asarray(tt, float),
            s["filters"], s["lc_data"], s["priors"], sys_plan=s["sys_plan"], detection_limit=s["limit"],
            z_table=s["z_table"])
    return _ORACLES[name]


def _cpu_eval(task):
    from oracle import harness
    name, chunk = task
    lik, fixed = _oracle_for(name)
    return harness.oracle_logl(lik, fixed, chunk, _SPECS[name]["cols"])


# The reference's OWN classes, when a copy of the package travels with the repository (baseline/_ref/nmma: `pip install
# --no-deps --target baseline/_ref /root/reference`, git-ignored).  Third-party modules that are absent
# offline are stubbed exactly as for the golden vectors (tests/golden/reference_stubs.py: Keras forward pass = fp32
# NumPy on the same weights, astropy Planck18 = the flat-LCDM integral); every line of nmma/em/{model,em_likelihood,
# systematics,utils,lightcurve_generation}.py and nmma/core/{base,conversion}.py runs unmodified.
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
_REFLIKS = {}


def reference_classes_available():
    return os.path.isfile(os.path.join(REF_DIR, "nmma", "em", "em_likelihood.py"))


def _reference_lik_for(name):
    import contextlib
    if name not in _REFLIKS:
        with contextlib.redirect_stdout(sys.stderr):   # the reference prints import-time notices: stdout carries ONE JSON line
            _REFLIKS[name] = _build_reference_lik(name)
    return _REFLIKS[name]


def _build_reference_lik(name):
    if True:
        import tempfile
        import joblib
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import reference_stubs as RS
        from oracle.nmma_oracle import KerasStandIn
        RS.REF = REF_DIR
        mods = RS.load_reference()
        s = _SPECS[name]
        assert s["kind"] == "mlp"
        filters = list(s["filters"])
        tmp = tempfile.mkdtemp(prefix="nmma_ref_model_")
        # the surrogate in the reference's on-disk layout: {model}.joblib (SVD basis per filter) + {model}_tf/{filter}.keras
        # (placeholder files: the stubbed keras.saving.load_model hands out the in-memory weights of the workload)
        meta = {f.replace(":", "_"): {k: v for k, v in s["core"][f].items() if k != "model"} for f in filters}
        joblib.dump(meta, os.path.join(tmp, f"{s['model']}.joblib"))
        os.makedirs(os.path.join(tmp, f"{s['model']}_tf"))
        weights = {}
        for f in filters:
            fn = os.path.join(tmp, f"{s['model']}_tf", f.replace(":", "_") + ".keras")
            open(fn, "w").close()
            weights[fn] = s["core"][f]["model"]
        sys.modules["keras.saving"].load_model = lambda fn, compile=False: KerasStandIn(*weights[fn])
        model = mods["model"].SVDLightCurveModel(s["model"], svd_path=tmp, interpolation_type="tensorflow",
                                                 filters=filters, local_only=True)
        handler = mods["systematics"].FilterSystematicsHandler(filters, None, 1.0, s["lc_data"][0])
        return mods["em_likelihood"].EMTransientLikelihood(model, s["lc_data"], handler, s["priors"],
                                                           filters=filters, detection_limit=s["limit"])


def _ref_eval(task):
    import contextlib
    name, chunk = task
    lik = _reference_lik_for(name)
    cols = _SPECS[name]["cols"]
    with contextlib.redirect_stdout(sys.stderr):
        return np.array([lik.log_likelihood(dict(zip(cols, map(float, row)))) for row in chunk])


class CpuArm:
    """multiprocessing pool over the host cores (the analogue of the reference's schwimmbad task farm,
    nmma/core/mpi_setup.py:651-683); workers are forked before CUDA is initialised."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or len(os.sched_getaffinity(0))
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init)

    def run(self, name, pts, parts=None, reference_classes=False):
        if len(pts) == 0:
            return np.zeros(0), 0.0
        chunks = np.array_split(pts, min(len(pts), parts or self.cores * 4))
        t0 = time.perf_counter()
        out = np.concatenate(self.pool.map(_ref_eval if reference_classes else _cpu_eval, [(name, c) for c in chunks]))
        return out, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


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
# reference arm
# ---------------------------------------------------------------------------------------------
REFERENCE_WHAT = ("the reference's own classes (nmma.em.model.SVDLightCurveModel + nmma.em.em_likelihood.EMTransientLikelihood."
                  "log_likelihood(dict) per point, unmodified source from baseline/_ref; absent third-party modules stubbed: "
                  "Keras forward pass = fp32 NumPy on the same weights, astropy Planck18 = flat-LCDM integral)")
PORT_WHAT = ("the per-point NumPy/SciPy oracle port of the reference (fp32 NumPy MLP standing in for Keras; baseline/_ref "
             "absent)")


def reference_arm(args):
    """The reference's own CPU implementation of the path on all host cores; each step is a bounded sample of the
    workload.  When the reference package travels with the repository (baseline/_ref/nmma, see REF_DIR) its OWN classes
    run, with the third-party modules that are absent offline (bilby, sncosmo, astropy, keras) stubbed as for the golden
    vectors; otherwise the oracle PORT (same NumPy / SciPy / sklearn calls), which reproduces the reference's source files
    to 1e-15 on the committed vectors (tests/test_reference_vectors.py).  The line says which one ran (`cpu_baseline.kind`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _SPECS["c2"] = make_spec("c2")
    spec = _SPECS["c2"]
    arm = CpuArm()
    refcls = reference_classes_available()
    rng = np.random.default_rng(1234)
    probe, _ = spec["priors"].sample_array(arm.cores * 40, rng, spec["cols"])
    arm.run("c2", probe, reference_classes=refcls)         # builds the likelihood in every worker
    _, dt = arm.run("c2", probe, reference_classes=refcls)
    rate = len(probe) / dt
    # bounded sample: ~4 s of CPU work per step, shrunk so that steps + warm-up stay within ~2 minutes whatever K is
    sec_per_step = min(4.0, max(0.2, 120.0 / max(1, args.steps + args.warmup)))
    per_step = int(max(arm.cores * 40, min(rate * sec_per_step, 200000)))
    pts, _ = spec["priors"].sample_array(per_step, rng, spec["cols"])
    for _ in range(args.warmup):
        arm.run("c2", pts, reference_classes=refcls)
    t = 0.0
    for _ in range(args.steps):
        _, dt = arm.run("c2", pts, reference_classes=refcls)
        t += dt
    arm.close()
    value = per_step * args.steps / t
    sample = (f"{per_step} prior draws per step through " + (REFERENCE_WHAT if refcls else PORT_WHAT) +
              ", multiprocessing pool on all host cores")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 MLP / f64 likelihood",
        "data": "synthetic weights, real AT2017gfo photometry",
        "config": {"workload": WORKLOAD, "points_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "reference" if refcls else "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def rel_err(got, ref):
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    if got.shape != ref.shape:
        return float("inf"), False
    sent = -1.7976931348623157e308
    masks_equal = bool(np.array_equal(got == sent, ref == sent))
    ok = ref != sent
    if not ok.any() or not masks_equal:
        return (0.0 if masks_equal else float("inf")), masks_equal
    return float((np.abs(got[ok] - ref[ok]) / np.maximum(1.0, np.abs(ref[ok]))).max()), masks_equal


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from nmma_b200.sharding import HostResultBuffer, ShardedEvaluator, shard_bounds
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    M = args.points
    wanted = [c for c in args.configs.split(",") if c]
    names = ["c2"] + [c for c in ("c3", "c4", "c5") if c in wanted]
    for n in names:
        _SPECS[n] = make_spec(n)
    sizes = {"c2": M, "c3": args.points_c3, "c4": args.points_c4}

    # ---- inputs: rank r draws its blocks with seed 1234 + r; the first N_PARITY rows of the LAST timed batch are the
    # parity rows (every rank can rebuild every other rank's parity rows, rank 0 scores them with the oracle) ----
    def batches_for(name, r, count, n):
        s = _SPECS[name]
        rng = np.random.default_rng(1234 + 1000 * list(_SPECS).index(name) + r)
        return [s["priors"].sample_array(n, rng, s["cols"])[0] for _ in range(count)]

    def parity_rows(name, r, n):
        s = _SPECS[name]
        rng = np.random.default_rng(777 + 1000 * list(_SPECS).index(name) + r)
        return s["priors"].sample_array(n, rng, s["cols"])[0]

    n_par = {"c2": N_PARITY, "c3": N_PARITY, "c4": 16, "c5": 64}

    # ---- CPU work first (fork before CUDA init), rank 0 only ----
    cpu, oracle_ref = None, {}
    if rank == 0:
        arm = CpuArm()
        if world == 1 and not args.no_cpu:
            s = _SPECS["c2"]
            rng = np.random.default_rng(99)
            refcls = reference_classes_available()
            probe, _ = s["priors"].sample_array(arm.cores * 40, rng, s["cols"])
            arm.run("c2", probe, reference_classes=refcls)                             # builds the likelihood in every worker
            _, dt = arm.run("c2", probe, reference_classes=refcls)
            n_s = int(max(arm.cores * 40, min(len(probe) / dt * 12.0, 400000)))      # ~12 s of CPU work
            cpu_pts, _ = s["priors"].sample_array(n_s, rng, s["cols"])
            _, dt = arm.run("c2", cpu_pts, reference_classes=refcls)
            t0 = time.perf_counter()
            arm.run("c2", cpu_pts[:300], parts=1, reference_classes=refcls)
            one_core = 300 / (time.perf_counter() - t0)
            cpu = {"value": n_s / dt, "unit": UNIT, "cores": arm.cores, "kind": "reference" if refcls else "port",
                   "sample": f"{n_s} draws of the same prior through " + (REFERENCE_WHAT if refcls else PORT_WHAT) +
                             ", multiprocessing pool on all host cores",
                   "one_core_value": one_core, "cpu_count": os.cpu_count()}
        for name in names:
            if name == "c5":
                continue
            pts = np.concatenate([parity_rows(name, r, n_par[name]) for r in range(world)])
            oracle_ref[name] = arm.run(name, pts)[0].reshape(world, -1)
        if "c5" in names:
            from oracle import philox as ph
            s = _SPECS["c5"]
            kinds, params, tables = s["priors"].device_plan(s["cols"])
            kind_names = {0: "Uniform", 1: "DeltaFunction", 2: "Sine", 3: "Cosine", 4: "Gaussian", 5: "TruncatedGaussian",
                          6: "PowerLaw", 7: "Triangular", 8: "Interped"}
            blocks = []
            for r in range(world):
                lo, hi = shard_bounds(args.sweep_total, world, r)
                first = lo + (hi - lo) // 3        # well inside the shard, not at its head
                u = ph.unit_cube(SWEEP_SEED, first, n_par["c5"], len(s["cols"]))
                blocks.append(np.stack([ph.rescale_column(kind_names[int(k)], params[j], u[:, j], tables.get(j))
                                        for j, k in enumerate(kinds)], axis=1))
            oracle_ref["c5"] = arm.run("c5", np.concatenate(blocks))[0].reshape(world, -1)
        arm.close()

    # nvidia-smi needs ~1 s to start: launch it now, keep the samples that fall inside the timed regions
    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.set_device(local_rank)
    numa = None
    if world > 1:      # after the CPU arm has forked its workers: staging buffers and copy threads on the GPU's own NUMA node
        from nmma_b200.sharding import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up (NCCL_DEBUG=VERSION in some
        # environments): stdout carries ONE JSON line, so file descriptor 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    sh = ShardedEvaluator(lambda p: None) if world > 1 else None

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass

    def run_config(name, steps, warmup, with_e2e):
        """Device-timed weak-scaled run of one configuration (+ optional e2e arm); returns a dict for the JSON line."""
        spec = _SPECS[name]
        cols, P = spec["cols"], len(spec["cols"])
        n = sizes[name]
        lik = gpu_likelihood(spec, local_rank)
        eng = lik.sub_model.engine_for(cols)
        if name == "c2" and args.path:
            eng.set_option("path", args.path)
        nrot = N_ROTATE if n * P * 8 * N_ROTATE > 130e6 else max(N_ROTATE, int(260e6 / (n * P * 8)) + 1)
        host = batches_for(name, rank, nrot, n)
        last = (warmup + steps - 1) % nrot
        host[last][:n_par[name]] = parity_rows(name, rank, n_par[name])
        host_t = [torch.from_numpy(b).pin_memory() for b in host]
        dev_t = [b.to(dev) for b in host_t]
        local_bufs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(2)]
        full_bufs = [torch.empty(n * world, dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else local_bufs

        def step_device(i):
            if world == 1:
                eng.logl_device(dev_t[i % nrot], out=local_bufs[i % 2])
                return i % 2
            return sh.gather_overlapped(lambda out: eng.logl_device(dev_t[i % nrot], out=out), local_bufs, full_bufs, i)

        for i in range(warmup):
            step_device(i)
        if sh:
            sh.drain()
        sync()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.get_info("launches")
        w0 = time.perf_counter()
        e0.record()
        slot = 0
        for i in range(steps):
            evs[i][0].record()
            slot = step_device(warmup + i)
            evs[i][1].record()       # on the compute stream: kernel time only, the overlapped gather is not inside
        if sh:
            sh.drain()
        e1.record()
        sync()
        w1 = time.perf_counter()
        launches = eng.get_info("launches") - l0
        timed_path = eng.get_info("last_path")      # the kernel family of the TIMED launches (later small calls take others)
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        # parity on the timed output: the parity rows head every rank's block of the last step's gathered vector
        parity = None
        got = full_bufs[slot].view(world, n)[:, :n_par[name]].cpu().numpy() if world > 1 else \
            local_bufs[slot][:n_par[name]].cpu().numpy()[None, :]
        if rank == 0:
            err, masks = rel_err(got, oracle_ref[name])
            parity = {"rows": int(got.size), "ranks_checked": world, "max_rel_err": err, "sentinel_masks_equal": masks,
                      "tolerance": 1e-4, "what": "rows of the last TIMED step's output (all ranks' blocks of the gathered vector)"}
            assert masks and err < 1e-4, f"{name}: GPU/oracle disagree on the timed output: {err}"
        res = {"ms_total": ms_total, "kern_ms": kern_ms, "launches": int(launches), "parity": parity, "window": (w0, w1),
               "n": n, "P": P, "eng": eng, "lik": lik, "nrot": nrot, "path": timed_path}
        if with_e2e:
            hb = None
            if world > 1:
                hb = HostResultBuffer(n * world, rank, world, name=f"nmma_b200_bench_{os.environ.get('MASTER_PORT', '0')}_{name}")
                dist.barrier()
                hb.attach()
                out_np = hb.local
            else:
                out_host = torch.empty(n, dtype=torch.float64).pin_memory()
                out_np = out_host.numpy()

            def step_e2e(i):
                lik.log_likelihood_batch(host_t[i % nrot].numpy(), cols, out=out_np)   # H2D + kernels + D2H, synchronised
                if world > 1:
                    dist.barrier()          # the host consumer (rank 0) sees the complete vector after every step

            for i in range(max(warmup, 3)):
                step_e2e(i)
            sync()
            t0 = time.perf_counter()
            for i in range(steps):
                step_e2e(warmup + i)
            sync()
            e2e_s = max_over_ranks(time.perf_counter() - t0)
            res["window"] = (w0, time.perf_counter())
            pe = None
            if rank == 0:
                full = hb.full.reshape(world, n)[:, :n_par[name]] if world > 1 else out_np[None, :n_par[name]]
                err, masks = rel_err(np.array(full), oracle_ref[name])
                pe = {"rows": int(full.size), "max_rel_err": err, "sentinel_masks_equal": masks,
                      "what": "rows of the host result vector after the last timed e2e step (all ranks' slices)"}
                assert masks and err < 1e-4, f"{name}: e2e output disagrees with the oracle: {err}"
            if hb is not None:
                dist.barrier()
                hb.close()
            res["e2e_s"], res["parity_e2e"] = e2e_s, pe
        return res

    def roofline_for(name, r):
        eng, n = r["eng"], r["n"]
        flop = eng.get_info("algorithmic_flop_per_eval")
        achieved = n * flop / (r["kern_ms"] * 1e-3) / 1e12
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_gbs = n * (r["P"] * 8 + 8) / (r["kern_ms"] * 1e-3) / 1e9
        out = {"algorithmic_flop_per_eval": flop, "kernel_ms_per_launch": r["kern_ms"],
               "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak, "frac": hbm_gbs / hbm_peak,
                       "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"}}
        path = r["path"]
        if _SPECS[name]["kind"] == "gp":
            out.update({"bound": "fp64 (CUDA cores: F K Ntr kernel values (1 + r^2 q)^-alpha per evaluation)",
                        "achieved": achieved, "peak": pk["dfma"], "unit": "TFLOP/s", "frac": achieved / pk["dfma"],
                        "peak_source": "fp64 FMA micro-benchmark measured in this run (nmma_b200_dfma_peak)",
                        "kernel": "fused_gp_logl_kernel" if path == 4 else "coeff_gp_kernel + backend_logl_kernel"})
        elif path == 3:
            executed = eng.get_info("tc_executed_flop_per_eval")
            ex_t = n * executed / (r["kern_ms"] * 1e-3) / 1e12
            out.update({"bound": "tensor", "achieved": achieved, "peak": pk["f16"], "unit": "TFLOP/s",
                        "frac": achieved / pk["f16"], "peak_source": pk["f16_source"],
                        "tf32_tflops_measured_in_run": pk["tf32"],
                        "executed_tensor_tflops": ex_t, "executed_frac_of_peak": ex_t / pk["f16"],
                        "executed_tensor_flop_per_eval": executed, "frac_of_ffma_peak": achieved / pk["ffma"],
                        "kernel": "fused_tc_logl_kernel (tcgen05 kind::f16, fp16 hi/lo split operands)"})
        else:
            out.update({"bound": "fp32_fma (CUDA cores)", "achieved": achieved, "peak": pk["ffma"], "unit": "TFLOP/s",
                        "frac": achieved / pk["ffma"], "peak_source": "FFMA micro-benchmark measured in this run",
                        "kernel": {1: "fused_mlp_logl_kernel (FFMA)", 2: "two-stage kernels"}.get(path, "?")})
        return out

    # ---- headline configuration ----
    main = run_config("c2", args.steps, args.warmup, with_e2e=True)
    eng = main["eng"]
    pk = {"ffma": max(eng.ffma_peak(0, 20000), eng.ffma_peak(1, 20000)) / 1e12, "dfma": eng.dfma_peak(20000) / 1e12,
          "tf32": eng.tf32_peak(20000) / 1e12,
          # the contraction runs on kind::f16 MMAs: the dense bf16 / fp16 rate of MEASURED_PEAKS.json (burst: the kernel is
          # timed alone, a few ms per launch) is the denominator; the recipe's fallback otherwise
          "f16": float(peaks.get("bf16_tflops", 1590.0)),
          "f16_source": ("MEASURED_PEAKS.json bf16_tflops (cuBLAS dense bf16 burst; kind::f16 runs at the same rate)" if "bf16_tflops" in peaks
                         else "fallback 1.59 PFLOP/s (B200_PROFILING.md)")}
    window = main["window"]

    # ---- the other BASELINE.json configurations (short runs) ----
    config_lines = {}
    for name in names[1:]:
        if name == "c5":
            continue
        r = run_config(name, args.steps_cfg, 3, with_e2e=False)
        if rank == 0:
            n = r["n"]
            config_lines[name] = {
                "workload": _SPECS[name]["workload"], "value": n * world * args.steps_cfg / (r["ms_total"] * 1e-3), "unit": UNIT,
                "scaling": "weak", "points_per_gpu_per_step": n, "steps": args.steps_cfg, "ms_per_step": r["ms_total"] / args.steps_cfg,
                "P": r["P"], "gpu_launches": r["launches"], "roofline": roofline_for(name, r), "parity_in_run": r["parity"],
                "l2": f"{r['nrot']} rotating input batches"}
        del r

    if "c5" in names:
        spec = _SPECS["c5"]
        cols = spec["cols"]
        lik5 = main["lik"]                       # same likelihood as c2: the sweep adds the on-device prior draws
        total = args.sweep_total
        lo, hi = shard_bounds(total, world, rank)
        out_local = torch.empty(hi - lo, dtype=torch.float64, device=dev)
        out_full = torch.empty(total, dtype=torch.float64, device=dev) if world > 1 else out_local
        eng5 = lik5.sub_model.engine_for(cols)

        def sweep_once():
            eng5.logl_sweep(hi - lo, seed=SWEEP_SEED, first_index=lo, out=out_local)
            if world > 1:
                sh.gather(out_local, n_global=total, out=out_full)

        sweep_once()                              # warm-up pass over a short prefix would not warm the allocator: full pass
        sync()
        l0 = eng5.get_info("launches")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sweep_once()
        e1.record()
        sync()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches5 = eng5.get_info("launches") - l0
        got = []
        for r_ in range(world):
            rlo, rhi = shard_bounds(total, world, r_)
            first = rlo + (rhi - rlo) // 3
            got.append(out_full[first:first + n_par["c5"]].cpu().numpy())
        finite = int(torch.isfinite(out_full).sum().item()) if rank == 0 else 0
        if rank == 0:
            err, masks = rel_err(np.stack(got), oracle_ref["c5"])
            assert masks and err < 1e-4, f"c5: sharded sweep disagrees with the oracle: {err}"
            rate = total / (ms * 1e-3)
            flop = eng5.get_info("algorithmic_flop_per_eval")
            config_lines["c5"] = {
                "workload": spec["workload"], "value": rate, "unit": UNIT, "scaling": "strong", "total_points": total,
                "ms_per_sweep": ms, "points_per_gpu": hi - lo, "gpu_launches": int(launches5),
                "collective": f"one ncclAllGather of logL[{total}] ({total * 8 / 1e6:.0f} MB) after the sweep" if world > 1 else "none (1 GPU)",
                "h2d_bytes": 0, "d2h_bytes": 0, "finite_rows": finite,
                "roofline": {"bound": "tensor", "achieved": rate / world * flop / 1e12, "peak": pk["f16"], "unit": "TFLOP/s",
                             "frac": rate / world * flop / 1e12 / pk["f16"], "per": "GPU, draws + gather inside the timed region",
                             "peak_source": pk["f16_source"]},
                "parity_in_run": {"rows": int(np.stack(got).size), "ranks_checked": world, "max_rel_err": err,
                                  "sentinel_masks_equal": masks, "tolerance": 1e-4,
                                  "what": "rows of the TIMED gathered sweep output at 1/3 of every rank's shard vs the oracle on "
                                          "the same Philox draws rebuilt on the CPU (oracle/philox.py)"}}

    # ---- one-point latency through the bilby seam (pymultinest-style sampler: one dict per call) ----
    latency = None
    if rank == 0:
        lik = main["lik"]
        cols = _SPECS["c2"]["cols"]
        row = parity_rows("c2", 0, 8)
        dicts = [dict(zip(cols, r_)) for r_ in row]
        for d in dicts:
            lik.log_likelihood(d)
        t0 = time.perf_counter()
        reps = 400
        for i in range(reps):
            lik.log_likelihood(dicts[i % 8])
        dt = (time.perf_counter() - t0) / reps
        latency = {"us_per_call": dt * 1e6, "evals_per_s": 1.0 / dt, "api": "EMTransientLikelihood.log_likelihood(dict), N = 1",
                   "cpu_one_core_evals_per_s": cpu["one_core_value"] if cpu else None}

    w2 = time.perf_counter()
    clocks = None
    if sampler:
        clocks = sampler.stop(window[0], window[1])
        if clocks is not None:
            clocks["window"] = "device-timed steps and e2e steps of the headline configuration (contiguous, GPU busy throughout)"

    if rank == 0:
        total_pts = M * world
        P = main["P"]
        value = total_pts * args.steps / (main["ms_total"] * 1e-3)
        roofline = roofline_for("c2", main)
        roofline["ffma_peak_measured_tflops"] = pk["ffma"]
        roofline["dfma_peak_measured_tflops"] = pk["dfma"]
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        roofline["traffic"] = traffic
        last_path = main["path"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_total"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 MLP (fp16 hi/lo split operands on tcgen05 kind::f16, fp32 accumulate) / f64 likelihood" if last_path == 3 else "f32 MLP (FFMA) / f64 likelihood",
            "data": "synthetic: random-init Bu2019lm-shaped weights, real AT2017gfo photometry",
            "config": {"workload": WORKLOAD, "points_per_gpu_per_step": M, "global_points_per_step": total_pts,
                       "sharding": (f"contiguous row blocks x{world} (ShardedEvaluator), NCCL all-gather of logL on a side stream "
                                    f"under the next step's kernels") if world > 1 else "single GPU",
                       "l2": f"inputs rotate through {N_ROTATE} distinct batches ({N_ROTATE * M * P * 8 / 1e6:.0f} MB > 126 MB L2)",
                       "kernel_path": roofline.get("kernel"),
                       "cpu_affinity": f"rank 0 bound to the CPUs of its GPU's NUMA node ({numa})" if numa else "unchanged"},
            "e2e": {"value": total_pts * args.steps / main["e2e_s"], "unit": UNIT,
                    "h2d_bytes_per_step": total_pts * P * 8, "d2h_bytes_per_step": total_pts * 8,
                    "api": "EMTransientLikelihood.log_likelihood_batch -> nmma_b200_logl_host (pinned host buffers)" +
                           ("; every rank's D2H lands in its slice of one page-locked shared-memory vector read by rank 0 "
                            "(HostResultBuffer), no collective" if world > 1 else ""),
                    "parity_in_run": main["parity_e2e"]},
            "gpu_launches": main["launches"],
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity_in_run": main["parity"],
            "config_lines": config_lines,
            "latency_one_point": latency,
            "wall_s": {"total": w2 - T_START},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


T_START = time.perf_counter()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--configs", type=str, default="c3,c4,c5", help="extra BASELINE.json configurations to run (config_lines)")
    ap.add_argument("--steps-cfg", dest="steps_cfg", type=int, default=10)
    ap.add_argument("--points-c3", dest="points_c3", type=int, default=1_000_000)
    ap.add_argument("--points-c4", dest="points_c4", type=int, default=200_000)
    ap.add_argument("--sweep-total", dest="sweep_total", type=int, default=SWEEP_TOTAL)
    ap.add_argument("--path", type=int, default=0)
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
