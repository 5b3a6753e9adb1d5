#!/bin/bash
# The measurements DESIGN.md / profiles/README.md cite for the final code of round 2 (1 x B200), in one GPU call.
OUT=gpurun_out/r02final_b; mkdir -p $OUT
timeout 200 python bench.py --steps 20 --warmup 5 > $OUT/bench_final_20steps.json 2> $OUT/bench_20.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_final_reference.json 2> $OUT/bench_ref.err
timeout 200 python bench.py --steps 200 --warmup 50 --no-cpu-baseline > $OUT/bench_final_200steps.json 2>/dev/null
timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline > $OUT/bench_final_1000steps.json 2>/dev/null
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --num-agents 8 --batch-envs 512 --use-ego-color > $OUT/bench_final_configs3_a8_b512.json 2>/dev/null
timeout 300 python scripts/raster_sweep.py $OUT/raster_sweep_final.json 1024 4096 16384 65536 > $OUT/raster_sweep.log 2>&1
timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid_final.txt 2>&1
timeout 100 python scripts/timeline.py 1024 100 900 > $OUT/timeline_late_final.txt 2>&1
timeout 100 python scripts/timeline.py 512 100 300 8 > $OUT/timeline_a8_final.txt 2>&1
timeout 100 python scripts/render_perf.py 1024 2 30 > $OUT/render_perf_final.txt 2>&1
timeout 200 python scripts/episode_profile.py > $OUT/episode_profile_final.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size
MCR_NO_GRAPH=1 timeout 300 ncu --clock-control none --metrics $M -s 260 -c 26 --csv --log-file $OUT/step_kernels_final.csv python scripts/profile_step.py 1024 40 > $OUT/step_kernels.log 2>&1
for f in bench_final_20steps bench_final_reference bench_final_200steps bench_final_1000steps bench_final_configs3_a8_b512; do python - $OUT/$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
    print(sys.argv[1].split("/")[-1], "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], (r.get("kernel_ms"), "frac %.3f" % r.get("frac", 0)) if r else "")
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
grep "step (events)" $OUT/timeline_*_final.txt; cat $OUT/render_perf_final.txt; tail -6 $OUT/episode_profile_final.txt | cut -c1-400; wc -l $OUT/launches_bench_final.csv $OUT/step_kernels_final.csv
