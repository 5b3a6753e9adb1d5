mkdir -p gpurun_out/r4p
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4p/smoke_ncu.csv python __graft_entry__.py --smoke > gpurun_out/r4p/smoke_ncu.log 2>&1; echo "ncu smoke rc=$?"; tail -2 gpurun_out/r4p/smoke_ncu.log; wc -l gpurun_out/r4p/smoke_ncu.csv
timeout 240 compute-sanitizer --tool memcheck python __graft_entry__.py --smoke > gpurun_out/r4p/smoke_memcheck.log 2>&1; echo "memcheck smoke rc=$?"; tail -4 gpurun_out/r4p/smoke_memcheck.log
timeout 240 compute-sanitizer --tool racecheck python __graft_entry__.py --smoke > gpurun_out/r4p/smoke_racecheck.log 2>&1; echo "racecheck smoke rc=$?"; tail -4 gpurun_out/r4p/smoke_racecheck.log
timeout 240 compute-sanitizer --tool synccheck python __graft_entry__.py --smoke > gpurun_out/r4p/smoke_synccheck.log 2>&1; echo "synccheck smoke rc=$?"; tail -4 gpurun_out/r4p/smoke_synccheck.log
