"""Summarise an ncu report: python scripts/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("launch__occupancy_limit_registers", "occ_reg"), ("launch__occupancy_limit_shared_mem", "occ_smem"), ("launch__occupancy_limit_warps", "occ_warps"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_act%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wr"),
        ("smsp__inst_executed.sum", "winst"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf")]
for r in data:
    name = r[hdr.index("Kernel Name")].split("(")[0][-30:]
    out = []
    for m, lab in want:
        if m in hdr:
            v = r[hdr.index(m)]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            out.append(f"{lab}={v}")
    print(name, " ".join(out))
