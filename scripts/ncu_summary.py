"""Prints the metrics we quote from an .ncu-rep (raw page) per captured launch; optional --source: hottest source lines."""
import csv
import subprocess
import sys

WANT = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, units = rows[0], rows[1]
for r in rows[2:]:
    print(f"## {r[H.index('Kernel Name')][:100]}")
    for w in WANT:
        if w in H:
            print(f"  {w:88s} {r[H.index(w)]} {units[H.index(w)]}")
if "--source" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    H = rows[0]
    col = H.index("# Instructions Executed") if "# Instructions Executed" in H else [i for i, h in enumerate(H) if "Instructions Executed" in h][0]
    s = H.index("Source")
    body = [r for r in rows[1:] if len(r) > col and r[col].replace(".", "").isdigit()]
    tot = sum(float(r[col]) for r in body)
    print("\nhottest source lines (share of executed instructions):")
    for r in sorted(body, key=lambda r: -float(r[col]))[:22]:
        print(f"  {100 * float(r[col]) / tot:5.1f}%  {r[s].strip()[:120]}")
