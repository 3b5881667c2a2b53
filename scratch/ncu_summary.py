"""Per-kernel summary of an `ncu --set full` report (raw page as CSV):
    ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python scratch/ncu_summary.py X_raw.csv"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
M = [("gpu__time_duration.sum", "time"),
     ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
     ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
     ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
     ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma_active_pct"),
     ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma_inst_pct"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
     ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster"),
     ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
     ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
     ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
     ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
     ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall_sleeping")]
ki = hdr.index("Kernel Name")
agg = defaultdict(list)
for r in data:
    agg[r[ki]].append(r)
for k, rs in agg.items():
    print(f"== {k}   ({len(rs)} captured launches)")
    for name, short in M:
        if name not in hdr:
            continue
        i = hdr.index(name)
        vals = []
        for r in rs:
            try:
                vals.append(float(r[i].replace(",", "")))
            except ValueError:
                pass
        if not vals:
            continue
        mean = sum(vals) / len(vals)
        print(f"   {short:18s} {mean:14.3f} {units[i]:16s} (min {min(vals):.3f} max {max(vals):.3f})   {name}")
