#!/bin/bash
# ncu source-level capture of the lane-cooperative tracker on a 1,024-stream bank
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:track_coop -s 60 -c 1 -f -o gpurun_out/coop \
    python tools/small_bank_bench.py 1024 > gpurun_out/coop_ncu.log 2>&1
tail -3 gpurun_out/coop_ncu.log
ls -la gpurun_out/coop.ncu-rep
