#!/bin/bash
# A/B of fill_kernel's frame store: 6 x STG.128 per thread vs shared-memory staging + one cp.async.bulk
# (UBLKCP) per frame.  Timing with CUDA events, then the store-path counters of one mid-episode launch of each.
OUT=gpurun_out/${1:-store_ab}; mkdir -p $OUT
M=gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_op_write.sum,lts__t_requests_op_write.sum,dram__bytes_write.sum,dram__bytes_read.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for V in stg bulk; do
  if [ $V = stg ]; then export MCR_LIB_PATH=multi_car_racing_b200/libmcr_stg.so; else unset MCR_LIB_PATH; fi
  echo "== $V"; python scripts/render_perf.py 1024 2 40 | tee $OUT/perf_$V.txt
  NCU="mid(170)" MCR_NO_GRAPH=1 timeout 600 ncu --clock-control none --profile-from-start off -k regex:fill_kernel --metrics $M --csv --log-file $OUT/ncu_$V.csv python scripts/render_perf.py 1024 2 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/ncu_$V.csv")) if len(r)>10]
hdr=rows[0]; im=hdr.index("Metric Name"); iv=hdr.index("Metric Value")
for r in rows[1:]: print("   %-70s %s" % (r[im], r[iv]))
PY
done
