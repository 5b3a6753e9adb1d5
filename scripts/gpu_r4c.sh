#!/bin/bash
OUT=gpurun_out/r4c; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid_$name.txt 2>&1
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench20_$name.json 2>$OUT/bench20_$name.err
  echo "== $name"; head -22 $OUT/timeline_mid_$name.txt | tail -21
  python - $OUT/bench20_$name.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("/")[-1], "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], r.get("kernel_ms"), "frac %.3f" % r.get("frac", 0))
    except Exception as e:
        print(f, "unreadable", e)
PY
}
run l2 A=1
run l2_noscoreprio MCR_SCORE_PRIO_OFF=1
run l1 MCR_FLAG_HANDOFF=1
run l0 MCR_FLAG_HANDOFF=0
timeout 300 python scripts/timeline.py 1024 100 900 > $OUT/timeline_late.txt 2>&1; echo "== late"; head -24 $OUT/timeline_late.txt
timeout 300 python scripts/timeline.py 512 100 300 8 > $OUT/timeline_a8.txt 2>&1; echo "== a8"; head -3 $OUT/timeline_a8.txt
timeout 300 python bench.py --batch-envs 4096 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench20_b4096.json 2>$OUT/bench20_b4096.err; echo "b4096 rc=$?"; cut -c1-200 $OUT/bench20_b4096.json
timeout 600 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline > $OUT/bench1000.json 2>$OUT/bench1000.err; echo "1000 rc=$?"; cut -c1-200 $OUT/bench1000.json
