#!/bin/bash
# multi-GPU session: probe at every N up to the visible GPUs, then bench.py at N (run through gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo_$N.txt
lscpu | grep -E "^CPU\(s\)|NUMA|Model name" > gpurun_out/cpu_$N.txt; cat gpurun_out/cpu_$N.txt
free -g | head -2
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  if [ $n -eq 1 ]; then PROBE_SECONDS=0.7 timeout 120 python tools/h2d_probe.py > gpurun_out/probe_${N}gpu_n$n.md 2>&1
  else PROBE_SECONDS=0.7 timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py > gpurun_out/probe_${N}gpu_n$n.md 2>&1; fi
  tail -6 gpurun_out/probe_${N}gpu_n$n.md
done
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_bench_$N.json 2> gpurun_out/scale_bench_$N.err
grep -E "nranks|NVLS" gpurun_out/scale_bench_$N.err | head -4
tail -3 gpurun_out/scale_bench_$N.err | cut -c1-400
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/scale_bench_$N.json') if l.startswith('{')][-1])
e=d['e2e']
print('N', d['n_gpus'], 'value %.0f ms/step %.2f'%(d['value'], d['ms_per_step']), d['collective'])
print('e2e %.0f  pcie/gpu %.1f  mode %s'%(e['value'], e['pcie_gbs'], e['h2d_mode']))
print('probe', {k:v for k,v in e['h2d_probe'].items() if k!='what'})
print('trials', e['h2d_mode_trials_ms'], 'frac_probe', e['frac_of_h2d_probe'], 'frac_ceiling', e['frac_of_h2d_ceiling'], 'numa', e['pinned_numa_node'])
print('lock', d['lock_stats'])
PY
