#!/bin/bash
# A/B of an alternative build against libmcr.so on the rasteriser alone and the whole step: bash scripts/gpu_ab_lib.sh <lib.so>
ALT=$PWD/multi_car_racing_b200/$1; OUT=gpurun_out/ab_$1; mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_parity_300 or handoff" > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
MCR_LIB_PATH=$ALT timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_parity_300 or handoff" > $OUT/pytest_alt.log 2>&1; tail -2 $OUT/pytest_alt.log
for i in 1 2; do
echo "== base"; timeout 100 python scripts/render_perf.py 1024 2 30 | head -3; timeout 100 python scripts/timeline.py 1024 200 60 | grep "step (events)"
echo "== alt"; MCR_LIB_PATH=$ALT timeout 100 python scripts/render_perf.py 1024 2 30 | head -3; MCR_LIB_PATH=$ALT timeout 100 python scripts/timeline.py 1024 200 60 | grep "step (events)"
done
