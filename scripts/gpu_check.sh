#!/bin/bash
OUT=gpurun_out/r5c; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; rc=$?; tail -3 $OUT/pytest.log
if [ $rc -ne 0 ]; then echo "pytest failed or hung rc=$rc"; tail -30 $OUT/pytest.log; exit 1; fi
timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid.txt 2>&1; head -28 $OUT/timeline_mid.txt | tr '\n' ';' | sed 's/  */ /g'; echo
MCR_ACTION_COPY=1 timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid_copy.txt 2>&1; grep "step (events)" $OUT/timeline_mid_copy.txt
timeout 150 python bench.py --steps 20 --warmup 5 > $OUT/bench20.json 2>$OUT/bench20.err; echo "bench20 rc=$?"
timeout 150 python bench.py --steps 200 --warmup 50 --no-cpu-baseline --no-e2e > $OUT/bench200.json 2>$OUT/bench200.err
timeout 200 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --no-e2e > $OUT/bench1000.json 2>$OUT/bench1000.err
python - $OUT/bench20.json $OUT/bench200.json $OUT/bench1000.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("/")[-1], "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), r.get("kernel_ms"), "frac %.3f" % r.get("frac", 0), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(f, "unreadable", e)
PY
