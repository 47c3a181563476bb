python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for p in ${PARTS:-0 2 3}; do
python bench.py --no-e2e --no-cpu --slab-parts $p 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; o=r['other_kernel']
print('parts',$p,'value %.0f Msym/s  ms/step %.2f | %s %.3f ms (fp32 %.2f) | %s %.3f ms (fp32 %.2f) | clocks %s'%(d['value'],d['ms_per_step'],r['kernel'],r['ms_per_launch'],r['fp32_issue']['frac'],o['kernel'],o['ms_per_launch'],o['fp32_issue']['frac'],d['clocks']))"
done
