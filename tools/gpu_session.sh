#!/bin/bash
# Typical B200 session (run through gpurun): parity tests, benchmark, then the ncu captures that
# tools/profile_summary.py turns into profiles/rNN_*.md / rNN_ncu_raw.csv.  Usage: bash tools/gpu_session.sh r02
set -x
R=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > gpurun_out/bench_ref_$R.json
python bench.py 2>gpurun_out/bench.err > gpurun_out/bench_$R.json
python tools/stage_bench.py 2>&1 | tail -14 > gpurun_out/stage_bench_$R.md
python tools/small_bank_bench.py 512 1024 2048 4096 8192 16384 32768 65536 > gpurun_out/small_bank_$R.md 2>&1
python tools/umma_bench.py > gpurun_out/umma_bench_$R.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'search_umma_batch_kernel|search_mma_batch_kernel' -s 8 -c 2 -o gpurun_out/prof_search_$R python tools/umma_bench.py > gpurun_out/ncu_search.log 2>&1
# launch list (every launch with its device time; compare shares)
CMD="python bench.py --streams 131072 --seconds 2 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$R.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu_list.log 2>&1
# full capture of the two RX kernels at full launch size (one slab)
CMD2="python bench.py --streams 131072 --seconds 1 --steps 1 --warmup 3 --no-e2e --no-cpu --slab-parts 1"
$CMD2 > gpurun_out/plain2_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'frontend_kernel|track_kernel' -s 10 -c 2 -o gpurun_out/prof_$R $CMD2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; o=r['other_kernel']
print('value %.0f Msym/s  ms/step %.2f | %s %.3f | %s %.3f'%(d['value'],d['ms_per_step'],r['kernel'],r['ms_per_launch'],o['kernel'],o['ms_per_launch']))"
# the opt-in fused tcgen05 front-end: A/B against the default, then one full capture (tools/ncu_extract.py reads it)
bash tools/fe_umma_ab.sh > gpurun_out/fe_umma_ab_$R.txt 2>&1
SC_FE_SEARCH=tcgen05 ncu --set full --clock-control none --import-source on -k regex:frontend_umma_kernel -s 10 -c 1 -o gpurun_out/prof_fe_umma_$R $CMD2 > gpurun_out/ncu_fe_umma.log 2>&1
