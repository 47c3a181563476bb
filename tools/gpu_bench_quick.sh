#!/bin/bash
# parity of the RX chain + a short device-resident bench line (for kernel iterations)
timeout 900 python -m pytest tests -q -m gpu -x -k "rx_gpu or stage_gpu or configs" 2>&1 | tail -3
python bench.py --no-e2e --no-cpu --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; o=r['other_kernel']
print('value %.0f Msym/s  ms/step %.2f | %s %.3f | %s %.3f'%(d['value'],d['ms_per_step'],r['kernel'],r['ms_per_launch'],o['kernel'],o['ms_per_launch']))"
