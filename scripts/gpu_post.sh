#!/bin/bash
OUT=gpurun_out/r4u; mkdir -p $OUT
timeout 90 python -m pytest tests/test_golden.py -m gpu -x -q > $OUT/pytest_golden.log 2>&1; rc=$?; tail -2 $OUT/pytest_golden.log
if [ $rc -ne 0 ]; then echo "golden failed or hung rc=$rc"; tail -30 $OUT/pytest_golden.log; exit 1; fi
timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid.txt 2>&1; rc=$?; head -28 $OUT/timeline_mid.txt | tr '\n' ';' | sed 's/  */ /g'; echo
if [ $rc -ne 0 ]; then echo "timeline failed or hung rc=$rc"; exit 1; fi
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; rc=$?; tail -3 $OUT/pytest.log
if [ $rc -ne 0 ]; then echo "pytest failed or hung rc=$rc"; tail -40 $OUT/pytest.log; exit 1; fi
MCR_LIB_PATH=$PWD/multi_car_racing_b200/libmcr_clk.so timeout 120 python scripts/post_phases.py 1024 2 2>&1 | tail -9
timeout 100 python scripts/timeline.py 1024 100 900 > $OUT/timeline_late.txt 2>&1; grep "step (events)" $OUT/timeline_late.txt
timeout 100 python scripts/timeline.py 512 100 300 8 > $OUT/timeline_a8.txt 2>&1; grep "step (events)" $OUT/timeline_a8.txt
timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench20.json 2>$OUT/bench20.err
timeout 150 python bench.py --steps 200 --warmup 50 --no-cpu-baseline --no-e2e > $OUT/bench200.json 2>$OUT/bench200.err
python - $OUT/bench20.json $OUT/bench200.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("/")[-1], "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], r.get("kernel_ms"), "frac %.3f" % r.get("frac", 0))
    except Exception as e:
        print(f, "unreadable", e)
PY
