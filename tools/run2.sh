set -x
python -m pytest tests/test_tx_gpu.py -x -q -m gpu 2>&1 | tail -30
python bench.py --streams 16384 --seconds 4 --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -5 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -5 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
