#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.md"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads active per instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction (instruction fetch)"),
    ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall imc_miss (constant cache)"),
]


def main():
    rep = sys.argv[1]
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]].split("(")[0].replace("void ", "") for r in data]
    print(f"source: `{rep}` (ncu --set full --clock-control none; per-launch, cold-ish cache, replayed)\n")
    print("| metric | unit | " + " | ".join(f"#{i} {n}" for i, n in enumerate(names)) + " |")
    print("|---|---|" + "---|" * len(names))
    for m, label in METRICS:
        if m not in col:
            continue
        i = col[m]
        print(f"| {label} (`{m}`) | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    # every pipe-utilisation / instruction-count / L1 wavefront column the capture holds
    import re
    extra = re.compile(r"(sm__inst_executed_pipe_.*pct_of_peak_sustained_active|sm__pipe_.*cycles_active.*pct_of_peak_sustained_active|smsp__inst_executed\.sum$|"
                       r"l1tex__data_pipe_lsu_wavefronts.*pct|l1tex__data_pipe_lsu_wavefronts\.sum$|l1tex__lsu_writeback_active.*pct|smsp__warps_eligible\.avg\.per_cycle_active|"
                       r"smsp__average_warp_latency_per_inst_issued|l1tex__t_sectors_pipe_lsu_mem_global_op_ld\.sum$|lts__t_sectors_srcunit_tex_op_read\.sum$)")
    for h in hdr:
        if extra.search(h) and h not in dict(METRICS):
            i = col[h]
            print(f"| `{h}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |")


if __name__ == "__main__":
    main()
