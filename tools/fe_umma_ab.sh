#!/bin/bash
# A/B of the fused front-end's search modes on one B200: parity test first, then the RX chain's device-timed bench line.
python -m pytest tests/test_round2_gpu.py -x -q -k "tensor_core_search" 2>&1 | tail -4
for mode in ${MODES:-direct tcgen05}; do
SC_FE_UMMA_DEBUG=1 SC_FE_SEARCH=$mode python bench.py --no-e2e --no-cpu 2>gpurun_out/fe_umma_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; o=r['other_kernel']
print('$mode: value %.0f Msym/s  ms/step %.2f | %s %.3f | %s %.3f'%(d['value'],d['ms_per_step'],r['kernel'],r['ms_per_launch'],o['kernel'],o['ms_per_launch']))"
done
grep -m2 "CTAs per SM" gpurun_out/fe_umma_err.txt
