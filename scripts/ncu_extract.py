#!/usr/bin/env python
"""Extract the judged metrics of one kernel from an .ncu-rep into a small CSV for profiles/.

    python scripts/ncu_extract.py gpurun_out/r1/ncu_wf_trace.ncu-rep profiles/ncu_k_wf_trace_r1.csv

Reads the report with `ncu -i ... --page raw --csv` (B200_PROFILING.md recipe) and keeps duration, DRAM bytes,
throughputs, occupancy, issue rate, warp efficiency, cache hit rates, local-memory traffic and the stall breakdown.
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sass__inst_executed_register_spilling", "memory_l2_theoretical_sectors_global", "memory_l2_theoretical_sectors_local",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(names)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "value", "unit"])
        w.writerow(["kernel", vals[col["Kernel Name"]], ""])
        for n in KEEP:
            if n in col:
                w.writerow([n, vals[col[n]], units[col[n]]])
        for n in names:
            if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
                v = vals[col[n]]
                try:
                    if float(v.replace(",", "")) >= 0.05:
                        w.writerow([n, v, units[col[n]]])
                except ValueError:
                    pass
    print(open(out).read())


if __name__ == "__main__":
    main()
