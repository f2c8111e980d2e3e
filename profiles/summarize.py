#!/usr/bin/env python
"""Summarise an .ncu-rep into the handful of numbers DESIGN.md / bench.py quote.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_uniform.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "smsp__cycles_active.avg",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]], rows[1], hdr


def main():
    path = sys.argv[1]
    recs, units, hdr = raw(path)
    unit = dict(zip(hdr, units))
    for rec in recs:
        print("== kernel:", rec.get("Kernel Name", "?")[:80])
        for k in KEYS:
            if k in rec:
                print(f"  {k:90s} {rec[k]:>18s} {unit.get(k, '')}")
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr = rows[0]
        try:
            i_src = hdr.index("Source"); i_samp = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Sampling Data (All)")
        except ValueError:
            print(hdr); return
        body = [r for r in rows[1:] if len(r) > i_samp]
        def val(r):
            try:
                return float(r[i_samp])
            except Exception:
                return 0.0
        tot = sum(val(r) for r in body) or 1.0
        print(f"== top {n} SASS lines by samples (total {tot:.0f})")
        for r in sorted(body, key=val, reverse=True)[:n]:
            print(f"  {100 * val(r) / tot:5.1f}%  {r[i_src][:110]}")


if __name__ == "__main__":
    main()
