#!/bin/bash
# One GPU call of the rasteriser loop: parity tests, render_kernel timing at three episode phases, per-phase
# cycles (instrumented build), ncu counters and the per-line profile of one mid-episode launch.
# usage: bash scripts/gpu_render_check.sh <tag> [pytest -k expression | "none"]
TAG=${1:-x}; KEXPR=${2:-}
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$KEXPR" != "none" ]; then
if [ -n "$KEXPR" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest.log 2>&1; else timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; fi
tail -3 $OUT/pytest.log
fi
python scripts/render_perf.py 1024 2 30 > $OUT/render_perf.txt 2>&1; cat $OUT/render_perf.txt
if [ -f multi_car_racing_b200/libmcr_clk.so ]; then MCR_LIB_PATH=multi_car_racing_b200/libmcr_clk.so python scripts/render_phases.py 1024 2 > $OUT/render_phases.txt 2>&1; cat $OUT/render_phases.txt; fi
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
NCU=all MCR_NO_GRAPH=1 timeout 600 ncu --clock-control none --profile-from-start off -k regex:"render_kernel|project_kernel|fill_kernel" --metrics $M --csv --log-file $OUT/ncu_render.csv python scripts/render_perf.py 1024 2 > $OUT/ncu_run.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/ncu_render.csv")) if len(r)>10]
hdr=rows[0]; im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
by={}
for r in rows[1:]:
    by.setdefault(r[iid],{})[r[im]]=r[iv]
for k,v in by.items():
    print(k, " ".join("%s=%s"%(a.split("__")[-1][:28],b) for a,b in v.items()))
PY
NCU="mid(170)" MCR_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"render_kernel|project_kernel|fill_kernel" -o $OUT/render_full -f python scripts/render_perf.py 1024 2 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/render_full.ncu-rep --page source --csv --print-source cuda,sass > $OUT/render_source.csv 2>/dev/null
python scripts/ncu_lines.py $OUT/render_source.csv 45 samples > $OUT/render_lines_by_samples.txt; head -50 $OUT/render_lines_by_samples.txt
python scripts/ncu_lines.py $OUT/render_source.csv 60 > $OUT/render_lines_by_inst.txt
ncu -i $OUT/render_full.ncu-rep --page raw --csv > $OUT/render_raw.csv 2>/dev/null
