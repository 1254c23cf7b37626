#!/usr/bin/env python3
"""Extracts the counters DESIGN.md / bench.py quote from the raw page of an ncu capture:
     ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python tools/ncu_raw_summary.py X_raw.csv "title" > summary.csv
Output: metric,unit,value lines (the format of profiles/r*_ncu_batch_exp_chunk_summary*.csv, which bench.py reads for
roofline.traffic)."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    title = sys.argv[2] if len(sys.argv) > 2 else "ncu --set full capture"
    hdr, units, vals = rows[0], rows[1], rows[2]
    print('metric,unit,"%s; kernel %s"' % (title.replace('"', "'"), vals[hdr.index("Kernel Name")].split("(")[0].replace('"', "'")))
    for w in WANT:
        hits = [i for i, h in enumerate(hdr) if h == w or h.endswith("." + w)]
        if hits:
            i = hits[0]
            print("%s,%s,%s" % (w, units[i], vals[i].replace(",", "")))


if __name__ == "__main__":
    main()
