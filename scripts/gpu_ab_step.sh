#!/bin/bash
# A/B of alternative builds on the whole step: bash scripts/gpu_ab_step.sh <lib.so> [<lib2.so> ...]
for i in 1 2; do
echo "== base"; timeout 100 python scripts/timeline.py 1024 200 60 | grep "step (events)\|sweep_end_packed\|  sweep  "
for L in "$@"; do echo "== $L"; MCR_LIB_PATH=$PWD/multi_car_racing_b200/$L timeout 100 python scripts/timeline.py 1024 200 60 | grep "step (events)\|sweep_end_packed\|  sweep  "; done
done
for L in "$@"; do MCR_LIB_PATH=$PWD/multi_car_racing_b200/$L timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_parity_300 or pose_after" 2>&1 | tail -1; done
