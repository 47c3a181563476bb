"""Per-instruction stall samples of track_coop_kernel from gpurun_out/coop.ncu-rep (ncu --page source)."""
import csv, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/coop.ncu-rep'
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = rows[1]; data = rows[2:]
ia = h.index('Address'); isrc = h.index('Source'); iall = h.index('Warp Stall Sampling (All Samples)'); iex = h.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[iall] or 0) for r in data)
print('total samples', tot)
agg = {h[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {k:28s} {v:6d} {v / tot:.1%}")
full = '--all' in sys.argv
for r in data:
    s = int(r[iall] or 0)
    if full or s >= tot * thr:
        top = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{int(r[ia], 16) & 0xffff:05x} {r[isrc][:64]:64s} {s:5d} {s / tot:6.1%} x{r[iex]:>6s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
