#!/bin/bash
# compute-sanitizer memcheck + racecheck over the kernels with shared-memory atomics and hand-placed barriers
# (VERDICT r1 next #8).  Logs -> gpurun_out/<tag>/, summaries are copied to profiles/ by hand.
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export MCR_NO_GRAPH=1
MEM_TESTS="tests/test_gpu_parity.py::test_reset_matches_oracle tests/test_gpu_parity.py::test_step_parity_300[2-4-3] tests/test_gpu_autoreset.py::test_fresh_tracks_follow_each_envs_own_stream[next_step-True] tests/test_gpu_parity.py::test_frame_stack_ring[4] tests/test_gpu_parity.py::test_fused_observation_formats[rgb_chw_f16] tests/test_gpu_parity.py::test_car_car_collisions"
RACE_TESTS="tests/test_gpu_parity.py::test_reset_matches_oracle tests/test_gpu_parity.py::test_step_parity_300[2-4-3] tests/test_gpu_parity.py::test_render_modes_between_steps"
(timeout 1500 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 9 python -m pytest -x -q $MEM_TESTS > $OUT/memcheck_tests.log 2>&1; echo "rc=$?" >> $OUT/memcheck_tests.log)
tail -4 $OUT/memcheck_tests.log
(timeout 1500 compute-sanitizer --tool racecheck --print-limit 30 --error-exitcode 9 python -m pytest -x -q $RACE_TESTS > $OUT/racecheck_tests.log 2>&1; echo "rc=$?" >> $OUT/racecheck_tests.log)
tail -4 $OUT/racecheck_tests.log
