"""Condense an .ncu-rep into the few metrics DESIGN.md / bench.py quote (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>.csv"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_shared.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "metric", "unit", "value"])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        for h, u, v in zip(hdr, units, r):
            if h in KEYS:
                w.writerow([name, h, u, v])
print("wrote", out)
