#!/bin/bash
OUT=gpurun_out/r5a; mkdir -p $OUT
timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/timeline_mid.txt 2>&1; rc=$?; head -28 $OUT/timeline_mid.txt | tr '\n' ';' | sed 's/  */ /g'; echo
if [ $rc -ne 0 ]; then echo "timeline failed or hung rc=$rc"; exit 1; fi
MCR_LIB_PATH=$PWD/multi_car_racing_b200/libmcr_clk.so timeout 120 python scripts/head_phases.py 1024 2 2>&1 | tail -8
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; rc=$?; tail -3 $OUT/pytest.log
if [ $rc -ne 0 ]; then echo "pytest failed or hung rc=$rc"; tail -40 $OUT/pytest.log; exit 1; fi
timeout 300 python scripts/raster_sweep.py $OUT/raster_sweep.json 4096 16384 65536 > $OUT/raster_sweep.log 2>&1
python - $OUT/raster_sweep.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
for r in d['rows']:
    print(r['batch_envs'], 'mid %.1f us frac %.3f | step %.3f ms %.4g af/s'%(r['render_ms_mid_episode']*1e3, r['frac_mid_episode'], r['step_ms'], r['step_agent_frames_per_s']))
PY
timeout 100 python scripts/timeline.py 512 100 300 8 2>&1 | grep "step (events)"
