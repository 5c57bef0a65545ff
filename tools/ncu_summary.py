#!/usr/bin/env python3
"""Summarise ncu captures (gpurun_out/*.ncu-rep) into profiles/<round>_<name>.txt: the handful of metrics the
roofline discussion in DESIGN.md uses, plus the per-source-line stall hot spots when -lineinfo mapped them."""
import csv
import io
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_blocks", "launch__occupancy_limit_warps",
        "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "smsp__inst_executed_op_shared_atom.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_sample_buffer_full"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return {}, ""
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    return d, d.get("Kernel Name", ("", ""))[0]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    reps = sys.argv[2:] or sorted(f for f in os.listdir("gpurun_out") if f.endswith(".ncu-rep"))
    os.makedirs("profiles", exist_ok=True)
    for rep in reps:
        path = rep if os.path.exists(rep) else os.path.join("gpurun_out", rep)
        d, name = raw(path)
        if not d:
            continue
        base = os.path.basename(path)[:-8]
        lines = ["ncu --set full --clock-control none, one launch: %s" % name, "source: gpurun_out/%s (not committed)" % os.path.basename(path), ""]
        for k in KEYS:
            if k in d:
                lines.append("%-75s %s %s" % (k, d[k][0], d[k][1]))
        extra = [k for k in d if ("stalled" in k and "per_warp_active" in k and k not in KEYS)]
        with open(os.path.join("profiles", "%s_ncu_%s.txt" % (rnd, base.replace("prof_", "").replace("final_", ""))), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("\n".join(lines[:1] + [l for l in lines[3:]]))
        print()


if __name__ == "__main__":
    main()
