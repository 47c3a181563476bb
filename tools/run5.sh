CMD="python bench.py --streams 131072 --seconds 1 --steps 1 --warmup 3 --no-e2e --no-cpu --slab-parts 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'frontend_kernel|track_kernel' -s 10 -c 2 -o gpurun_out/prof_r01b $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
