#!/usr/bin/env python
"""Compact text summary of an .ncu-rep (read here on the CPU box with `ncu -i`): the handful of counters that
justify the design choices (issue / pipe utilisation, occupancy, DRAM traffic, stall reasons)."""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_active.max",
    "smsp__cycles_active.avg", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "sm__sass_thread_inst_executed_op_fmul_pred_on.sum", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
]


def main(rep, title=""):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {title or rep}")
    print("# source: ncu --set full --clock-control none --import-source on (under gpurun); summary by tools/ncu_summary.py")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n## kernel: {d.get('Kernel Name', '?')}   grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for k in KEEP:
            if k in d:
                print(f"{k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
        stalls = [(float(d[h] or 0), h) for h in hdr
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        if not stalls:
            stalls = [(float(d[h] or 0), h) for h in hdr
                      if "issue_stalled" in h and h.endswith("per_warp_active.pct")]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"{h:75s} {v:18.3f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
