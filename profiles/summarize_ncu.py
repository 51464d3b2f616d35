#!/usr/bin/env python
"""Condense an `ncu --set full` report into the handful of numbers DESIGN.md and bench.py quote.

    python profiles/summarize_ncu.py gpurun_out/r01_fused.ncu-rep profiles/r01_fused_summary.json [--traffic]

Reads the report with `ncu -i ... --page raw --csv` (no GPU needed) and writes one JSON object per
profiled launch.  With --traffic it also refreshes profiles/traffic.json, which bench.py copies into
`roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch).
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "launch__occupancy_limit_registers": "occ_limit_regs_blocks",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_cycles_active_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "inst_pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "inst_pipe_alu_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "inst_pipe_fp64_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "inst_pipe_lsu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "inst_pipe_xu_pct",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active": "inst_pipe_tc_pct",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active": "inst_pipe_tmem_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_cycles_active_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_hmma_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_elapsed.max": "sm_cycles_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum": "thread_ffma",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "thread_dfma",
}
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
              "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return v


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    result = []
    for r in rows[2:]:
        rec = {"kernel": r[hdr.index("Kernel Name")]}
        stalls = {}
        for i, k in enumerate(hdr):
            if k in KEYS:
                v = num(r[i])
                if isinstance(v, float) and units[i] in UNIT_SCALE:
                    v *= UNIT_SCALE[units[i]]
                    unit = "s" if "second" in units[i] or units[i] in ("ms", "us", "ns", "s") else "B"
                    rec[KEYS[k] + "_" + unit] = v
                else:
                    rec[KEYS[k]] = v
            elif k.startswith(STALL_PREFIX) and k.endswith(STALL_SUFFIX) and "not_issued" not in k:
                v = num(r[i])
                if isinstance(v, float) and v >= 0.05:
                    stalls[k[len(STALL_PREFIX):-len(STALL_SUFFIX)]] = round(v, 3)
        rec["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        if "dram_read_B" in rec and "dram_write_B" in rec:
            rec["dram_bytes_per_launch"] = rec["dram_read_B"] + rec["dram_write_B"]
        result.append(rec)
    with open(out, "w") as fh:
        json.dump(result, fh, indent=1)
    print(json.dumps(result, indent=1))
    if "--traffic" in sys.argv and result:
        last = result[-1]
        with open(os.path.join(os.path.dirname(os.path.abspath(out)), "traffic.json"), "w") as fh:
            json.dump({"kernel": last["kernel"], "dram_bytes_per_launch": last.get("dram_bytes_per_launch"),
                       "source": os.path.basename(rep)}, fh, indent=1)


if __name__ == "__main__":
    main()
