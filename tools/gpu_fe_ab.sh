#!/bin/bash
# A/B of the fused front-end's search: parity suite under SC_FE_SEARCH=$1 (direct | mma), then bench of both
mkdir -p gpurun_out
SC_FE_SEARCH=${1:-mma} timeout 900 python -m pytest tests -q -m gpu -x -k "rx_gpu or configs or round2 or files or packet or dropin" 2>&1 | tail -4
for mode in direct mma; do
  SC_FE_SEARCH=$mode python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; o=r['other_kernel']
print('$mode value %.0f Msym/s  ms/step %.2f | %s %.3f | %s %.3f'%(d['value'],d['ms_per_step'],r['kernel'],r['ms_per_launch'],o['kernel'],o['ms_per_launch']))"
done
