set -x
CMD="python bench.py --streams 131072 --seconds 2 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_r01.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_list.log 2>&1
CMD2="python bench.py --streams 131072 --seconds 1 --steps 1 --warmup 3 --no-e2e --no-cpu --slab-parts 1"
$CMD2 > gpurun_out/plain2_r01.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'frontend_kernel|track_kernel' -s 10 -c 2 -o gpurun_out/prof_r01 $CMD2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
