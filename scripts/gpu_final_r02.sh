#!/bin/bash
# The measurements DESIGN.md / profiles/README.md cite for round 2, in one GPU call (1 x B200).
OUT=gpurun_out/r02final; mkdir -p $OUT
python bench.py --steps 20 --warmup 5 > $OUT/bench_20steps.json 2> $OUT/bench_20.err
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_ref.err
python bench.py --steps 200 --warmup 50 --no-cpu-baseline > $OUT/bench_200steps.json 2>/dev/null
python bench.py --steps 1000 --warmup 50 --no-cpu-baseline > $OUT/bench_1000steps.json 2>/dev/null
python scripts/raster_sweep.py $OUT/raster_sweep.json 1024 4096 16384 65536 > $OUT/raster_sweep.log 2>&1
python scripts/time_reset.py 1024 8192 65536 > $OUT/time_reset.txt 2>&1
python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid.txt 2>&1
python scripts/timeline.py 1024 100 900 > $OUT/timeline_late.txt 2>&1
python scripts/timeline.py 512 100 300 8 > $OUT/timeline_a8.txt 2>&1
python scripts/render_perf.py 1024 2 30 > $OUT/render_perf.txt 2>&1
MCR_RENDER_FUSED=1 python scripts/render_perf.py 1024 2 30 > $OUT/render_perf_fused.txt 2>&1
python scripts/episode_profile.py > $OUT/episode_profile.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
for f in bench_20steps bench_reference bench_200steps bench_1000steps; do python - $OUT/$f.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
print(sys.argv[1].split("/")[-1], "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], (r.get("kernel_ms"), "frac %.3f" % r.get("frac", 0)) if r else "")
PY
done
cat $OUT/time_reset.txt; grep step $OUT/timeline_*.txt; cat $OUT/render_perf.txt $OUT/render_perf_fused.txt; tail -12 $OUT/episode_profile.txt
