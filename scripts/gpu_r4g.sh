#!/bin/bash
OUT=gpurun_out/r4g; mkdir -p $OUT
timeout 90 python -m pytest tests/test_golden.py -m gpu -x -q > $OUT/pytest_golden.log 2>&1; rc=$?; tail -3 $OUT/pytest_golden.log
if [ $rc -ne 0 ]; then echo "golden failed or hung rc=$rc"; exit 1; fi
timeout 120 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid_l2.txt 2>&1; rc=$?; head -28 $OUT/timeline_mid_l2.txt
if [ $rc -ne 0 ]; then echo "timeline failed or hung rc=$rc"; exit 1; fi
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; rc=$?; tail -3 $OUT/pytest.log
if [ $rc -ne 0 ]; then echo "pytest failed or hung rc=$rc"; exit 1; fi
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench20_l2.json 2>$OUT/bench20_l2.err; echo "bench20 rc=$?"
MCR_FLAG_HANDOFF=0 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench20_l0.json 2>$OUT/bench20_l0.err
python - $OUT/bench20_l2.json $OUT/bench20_l0.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("/")[-1], "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), r.get("kernel_ms"), "frac %.3f" % r.get("frac", 0))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 120 python scripts/timeline.py 1024 100 900 > $OUT/timeline_late.txt 2>&1; echo "== late rc=$?"; head -30 $OUT/timeline_late.txt
