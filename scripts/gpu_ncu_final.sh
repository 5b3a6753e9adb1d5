#!/bin/bash
# ncu --set full of one mid-episode launch pair of the step-path rasteriser (final code): raw metrics for profiles/
OUT=gpurun_out/r5e; mkdir -p $OUT
NCU="mid(170)" MCR_NO_GRAPH=1 timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"project_kernel|fill_kernel" -o $OUT/render_full -f python scripts/render_perf.py 1024 2 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/render_full.ncu-rep --page raw --csv > $OUT/render_raw.csv 2>/dev/null
python - $OUT/render_raw.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_write.sum", "lts__t_requests_op_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:40])
    for i in idx[1:]:
        print("  %-90s %s" % (hdr[i], r[i]))
PY
