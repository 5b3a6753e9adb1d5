#!/bin/bash
# Refresh of the final bench lines and timelines (1 x B200) after the last kernel changes of round 2.
OUT=gpurun_out/r02final_c; mkdir -p $OUT
timeout 200 python bench.py --steps 20 --warmup 5 > $OUT/bench_final_20steps.json 2> $OUT/bench_20.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_final_reference.json 2> $OUT/bench_ref.err
timeout 200 python bench.py --steps 200 --warmup 50 --no-cpu-baseline > $OUT/bench_final_200steps.json 2>/dev/null
timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline > $OUT/bench_final_1000steps.json 2>/dev/null
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --num-agents 8 --batch-envs 512 --use-ego-color > $OUT/bench_final_configs3_a8_b512.json 2>/dev/null
timeout 300 python scripts/raster_sweep.py $OUT/raster_sweep_final.json 1024 4096 16384 65536 > $OUT/raster_sweep.log 2>&1
timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid_final.txt 2>&1
timeout 100 python scripts/timeline.py 1024 100 900 > $OUT/timeline_late_final.txt 2>&1
timeout 100 python scripts/timeline.py 512 100 300 8 > $OUT/timeline_a8_final.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
for f in bench_final_20steps bench_final_reference bench_final_200steps bench_final_1000steps bench_final_configs3_a8_b512; do python - $OUT/$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
    print(sys.argv[1].split("/")[-1], "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], (r.get("kernel_ms"), "frac %.3f" % r.get("frac", 0)) if r else "")
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
grep "step (events)" $OUT/timeline_*_final.txt
python - $OUT/raster_sweep_final.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
for r in d['rows']:
    print(r['batch_envs'], 'mid %.1f us frac %.3f | t0.02 %.1f us | step %.3f ms %.4g af/s'%(r['render_ms_mid_episode']*1e3, r['frac_mid_episode'], r['render_ms_t0.02']*1e3, r['step_ms'], r['step_agent_frames_per_s']))
PY
