#!/bin/bash
OUT=gpurun_out/r4s; mkdir -p $OUT
timeout 300 python scripts/raster_sweep.py $OUT/raster_sweep.json 1024 4096 16384 65536 > $OUT/raster_sweep.log 2>&1
python - $OUT/raster_sweep.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
for r in d['rows']:
    print(r['batch_envs'], 'mid %.1f us %.0f GB/s frac %.3f | t0.02 %.1f us frac %.3f | step %.3f ms %.4g af/s'%(r['render_ms_mid_episode']*1e3, r['render_GBs_mid_episode'], r['frac_mid_episode'], r['render_ms_t0.02']*1e3, r['frac_t0.02'], r['step_ms'], r['step_agent_frames_per_s']))
PY
timeout 200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_autoreset.py -m gpu -x -q 2>&1 | tail -2
timeout 100 python scripts/timeline.py 1024 200 60 2>&1 | grep "step (events)"
