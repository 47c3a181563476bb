set -x
CMD="python bench.py --streams 32768 --seconds 2 --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'frontend_kernel|track_kernel' -s 8 -c 4 -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
