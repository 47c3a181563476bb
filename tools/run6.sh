python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/ref.err | tee gpurun_out/bench_ref_r01.json | cut -c1-600
python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench_r01.json | cut -c1-3000
tail -3 gpurun_out/bench.err
