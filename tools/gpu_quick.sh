#!/bin/bash
# quick GPU iteration: selected tests (-k "$1") + stage bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "$1" 2>&1 | tail -25
timeout 600 python tools/stage_bench.py 2>&1 | tail -10 | tee gpurun_out/stage_quick.md
