#!/bin/bash
# flag hand-off A/B: parity tests, timelines and bench with the new library and the previous one (libmcr_old.so)
OUT=gpurun_out/r4a; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
for v in new old; do
  if [ $v = old ]; then export MCR_LIB_PATH=$PWD/multi_car_racing_b200/libmcr_old.so; else unset MCR_LIB_PATH; fi
  timeout 300 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid_$v.txt 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench20_$v.json 2>$OUT/bench20_$v.err
  timeout 300 python bench.py --steps 200 --warmup 50 --no-cpu-baseline --no-e2e > $OUT/bench200_$v.json 2>$OUT/bench200_$v.err
  echo "== $v"; grep "step (events)" $OUT/timeline_mid_$v.txt; head -22 $OUT/timeline_mid_$v.txt | tail -20
  python - $OUT/bench20_$v.json $OUT/bench200_$v.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("/")[-1], "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), r.get("kernel_ms"), "frac %.3f" % r.get("frac", 0))
    except Exception as e:
        print(f, "unreadable", e)
PY
done
unset MCR_LIB_PATH
MCR_NO_FLAG_HANDOFF=1 timeout 300 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid_nohandoff.txt 2>&1; echo "== new, MCR_NO_FLAG_HANDOFF=1"; grep "step (events)" $OUT/timeline_mid_nohandoff.txt
