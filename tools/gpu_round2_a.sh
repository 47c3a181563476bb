#!/bin/bash
# round 2, call A: parity tests, probe, bench, stage bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv | tail -1
lscpu | grep -E "Model name|^CPU\(s\)|NUMA" 
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/a_pytest.txt; tail -5 gpurun_out/a_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 120 python tools/h2d_probe.py > gpurun_out/a_probe1.md 2>&1; cat gpurun_out/a_probe1.md
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -c 600 gpurun_out/a_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/a_bench.json').read().strip().splitlines()[-1])
r=d['roofline']; o=r['other_kernel']; e=d['e2e']
print('value %.0f ms/step %.2f | %s %.3f ms | %s %.3f ms'%(d['value'],d['ms_per_step'],r['kernel'],r['ms_per_launch'],o['kernel'],o['ms_per_launch']))
print('e2e %.0f  pcie %.1f  probe %s  mode %s  trials %s frac %s'%(e['value'],e['pcie_gbs'],e['h2d_probe'],e['h2d_mode'],e['h2d_mode_trials_ms'],e['frac_of_h2d_probe']))
print('cpu', d['cpu_baseline'], d['cpu_baseline_fast'])
PY
timeout 600 python tools/stage_bench.py 2>&1 | tail -9 > gpurun_out/a_stage.md; cat gpurun_out/a_stage.md
