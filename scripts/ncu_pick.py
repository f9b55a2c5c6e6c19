"""stdin: `ncu --page raw --csv`; prints the metrics the round summaries quote, one block per captured launch."""
import csv, sys
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "sm__sass_inst_executed_op_local_ld.sum", "sm__sass_inst_executed_op_local_st.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    print("no data"); sys.exit(0)
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEYS:
        for h in hdr:
            if h == k or h.startswith(k):
                print(f"{h}: {d.get(h)}")
                break
    print("---")
