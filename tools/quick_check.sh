python -m pytest tests -x -q -m gpu 2>&1 | tail -1
python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; o=r['other_kernel']
print('value %.0f Msym/s  ms/step %.2f | %s %.3f | %s %.3f'%(d['value'],d['ms_per_step'],r['kernel'],r['ms_per_launch'],o['kernel'],o['ms_per_launch']))"
