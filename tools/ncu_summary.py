#!/usr/bin/env python3
"""Compact summary of an ncu --set full report: the metrics DESIGN.md / bench.py quote, per captured launch.

    python tools/ncu_summary.py gpurun_out/prof_step_X.ncu-rep > profiles/rNN_ncu_full_step_kernel.csv
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    w = csv.writer(sys.stdout)
    w.writerow(["launch", "kernel", "metric", "unit", "value"])
    for n, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")]
        for i, k in enumerate(hdr):
            if k in KEEP or (k.startswith(STALL) and k.endswith("_per_issue_active.ratio")):
                w.writerow([n, name, k, units[i], r[i]])


def traffic(report, envs, out_path):
    """profiles/traffic.json: DRAM bytes per env-step of the captured step-kernel launches (mean over launches)."""
    import json
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = []
    for r in rows[2:]:
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            b += float(r[i]) * scale[units[i]]
        tot.append(b)
    per_env = sum(tot) / len(tot) / envs
    json.dump({"kernel": rows[2][hdr.index("Kernel Name")], "launches": len(tot), "envs_per_launch": envs,
               "dram_bytes_per_env_step": per_env,
               "source": f"ncu --set full dram__bytes_read.sum + dram__bytes_write.sum, {os.path.basename(report)}"},
              open(out_path, "w"), indent=1)
    print(f"{per_env:.2f} B/env-step over {len(tot)} launches -> {out_path}", file=sys.stderr)


if __name__ == "__main__":
    import os
    if len(sys.argv) >= 5 and sys.argv[2] == "--traffic":
        traffic(sys.argv[1], int(sys.argv[3]), sys.argv[4])
    else:
        main()
