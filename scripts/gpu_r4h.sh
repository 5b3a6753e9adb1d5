#!/bin/bash
OUT=gpurun_out/r4h; mkdir -p $OUT
run() { # name, env...
  name=$1; shift
  env "$@" timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid_$name.txt 2>&1 || { echo "$name timeline failed/hung"; return; }
  env "$@" timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench20_$name.json 2>$OUT/bench20_$name.err
  echo "== $name"; head -28 $OUT/timeline_mid_$name.txt | tail -27 | tr '\n' ';' | sed 's/  */ /g'; echo
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
run s74 A=1
run s48 MCR_SCORE_CTAS=48
run s148 MCR_SCORE_CTAS=148
run classic MCR_SCORE_CLASSIC=1
run l0 MCR_FLAG_HANDOFF=0
