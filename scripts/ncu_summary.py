#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small JSON + markdown table.

usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ode_kernel [--note "..."]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__warps_active.avg.per_cycle_active": "warps_active_per_scheduler",
    "smsp__warps_eligible.avg.per_cycle_active": "warps_eligible_per_scheduler",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_utilisation_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct_of_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_active_pct_of_elapsed",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "thread_dfma",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "thread_dmul",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "thread_dadd",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed": "thread_dfma_per_cycle",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed": "thread_dmul_per_cycle",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed": "thread_dadd_per_cycle",
    "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained": "thread_dfma_peak_per_cycle",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instruction",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
    "smsp__sass_inst_executed_op_local_ld.sum": "local_load_instructions",
    "smsp__sass_inst_executed_op_local_st.sum": "local_store_instructions",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct": "local_load_l1_hit_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio": "stall_dispatch",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    launches = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS and vals[i] != "":
                try:
                    d[KEYS[h]] = float(vals[i].replace(",", ""))
                except ValueError:
                    d[KEYS[h]] = vals[i]
                if units[i]:
                    d[KEYS[h] + "_unit"] = units[i]
        if "thread_dfma_per_cycle" in d and "thread_dfma_peak_per_cycle" in d:
            flop_rate = 2 * d["thread_dfma_per_cycle"] + d.get("thread_dmul_per_cycle", 0) + d.get("thread_dadd_per_cycle", 0)
            d["fp64_flop_fraction_of_peak"] = flop_rate / (2 * d["thread_dfma_peak_per_cycle"])
        launches.append(d)
    json.dump({"source": rep, "note": note, "launches": launches}, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary: {rep}\n\n{note}\n\n")
        for d in launches:
            f.write(f"## {d['kernel']}\n\n| metric | value |\n|---|---|\n")
            for k, v in d.items():
                if k != "kernel" and not k.endswith("_unit"):
                    u = d.get(k + "_unit", "")
                    f.write(f"| {k} | {v:.6g} {u} |\n" if isinstance(v, float) else f"| {k} | {v} {u} |\n")
            f.write("\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
