"""Compact summary of an .ncu-rep (one kernel per line block). Usage:
    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [pairs_per_launch] > profiles/xxx.txt
"""
import csv
import io
import subprocess
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
import _ncu_pages  # noqa: E402
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
    "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_op_branch.sum", "smsp__sass_branch_targets_threads_divergent.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    pairs = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = _ncu_pages.page(rep, "raw")
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        unit = dict(zip(hdr, units))
        print("kernel:", rec.get("Kernel Name"))
        for k in KEYS:
            if k in rec:
                print(f"  {k} = {rec[k]} {unit[k]}")
        try:
            inst = float(rec["smsp__inst_executed.sum"])
            cyc = float(rec["sm__cycles_active.avg"])
            fp64 = float(rec["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]) / 100
            sms = 148
            fp64_warp = fp64 * cyc * 2 * sms  # FP64 pipe peak: 2 warp-inst / cycle / SM (64 lanes)
            print(f"  derived: FP64-pipe warp instructions ~= {fp64_warp:.4e}")
            if pairs:
                print(f"  derived: pairs per launch = {pairs:.4e}")
                print(f"  derived: thread instructions per pair (all pipes) = {inst * 32 / pairs:.1f}")
                print(f"  derived: FP64 thread instructions per pair = {fp64_warp * 32 / pairs:.1f}")
                dur = float(rec["gpu__time_duration.sum"])
                scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[unit["gpu__time_duration.sum"]]
                print(f"  derived: pair evals / s under ncu = {pairs / (dur * scale):.4e}")
        except (KeyError, ValueError):
            pass
        print()


if __name__ == "__main__":
    main()
