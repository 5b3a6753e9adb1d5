#!/bin/bash
OUT=gpurun_out/r4i; mkdir -p $OUT
timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/tl_copy.txt 2>&1; grep "step (events)" $OUT/tl_copy.txt
ACTION_INPLACE=1 timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/tl_inplace.txt 2>&1; grep "step (events)" $OUT/tl_inplace.txt
timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/tl_copy2.txt 2>&1; grep "step (events)" $OUT/tl_copy2.txt
ACTION_INPLACE=1 timeout 100 python scripts/timeline.py 1024 200 60 > $OUT/tl_inplace2.txt 2>&1; grep "step (events)" $OUT/tl_inplace2.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "handoff" 2>&1 | tail -3
