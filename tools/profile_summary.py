#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the committed summaries under profiles/.

usage: python tools/profile_summary.py ROUND   (reads gpurun_out/launches_rNN.csv and gpurun_out/prof_rNN.ncu-rep)
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel totals and shares ---------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"launches_{rnd}.csv"))) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.OrderedDict()
for r in rows:
    name = r[ki].split("(")[0].replace("void ", "").strip()
    a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
total = sum(a[1] for a in agg.values())
with open(os.path.join(out_dir, f"{rnd}_launches.md"), "w") as f:
    f.write(f"# {rnd}: every kernel launch of `python bench.py --streams 131072 --seconds 2 --steps 2 --warmup 3 --no-e2e --no-cpu`\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600` (cold-cache, serialised: compare SHARES).\n"
            "Launches include the synthesis of the input (tx_*), 5 RX passes x 4 calls x 2 slabs, the single-slab\n"
            "profiling pass and the statistics kernel.\n\n| kernel | launches | total us | share | grid (last) | block |\n|---|---|---|---|---|---|\n")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{name[:70]}` | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / total:.1f}% | {a[2]} | {a[3]} |\n")
    rx = {k: v for k, v in agg.items() if "frontend_kernel" in k or "track_kernel" in k}
    rxt = sum(v[1] for v in rx.values())
    f.write("\nShare inside the RX step (front-end + tracking only):\n\n")
    for k, v in rx.items():
        f.write(f"* `{k[:60]}`: {100 * v[1] / rxt:.1f}% ({v[1] / v[0] / 1e3:.1f} us per launch on average)\n")
with open(os.path.join(out_dir, f"{rnd}_launches.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "duration_ns"])
    for r in rows:
        w.writerow([r[0], r[ki].split("(")[0].replace("void ", "").strip()[:80], r[gi], r[bi], r[vi]])

# ---- full capture: selected metrics per kernel ------------------------------------------------------------
rep = os.path.join(ROOT, "gpurun_out", f"prof_{rnd}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units, data = rr[0], rr[1], rr[2:]
idx = {k: i for i, k in enumerate(h)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
stall = [k for k in h if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]

# compact, committed copy of the raw page: what bench.py's roofline.traffic is parsed from
git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
with open(os.path.join(out_dir, f"{rnd}_ncu_raw.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["# ncu --set full --clock-control none --import-source on; ncu -i prof.ncu-rep --page raw --csv; git", git])
    w.writerow(["kernel", "grid", "block", "metric", "unit", "value"])
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        for k in want + stall:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                w.writerow([name, r[idx["Grid Size"]], r[idx["Block Size"]], k, units[idx[k]], r[idx[k]].replace(",", "")])
with open(os.path.join(out_dir, f"{rnd}_ncu_summary.md"), "w") as f:
    f.write(f"# {rnd}: `ncu --set full --clock-control none` of the two RX kernels at full launch size\n\n"
            "Command: `python bench.py --streams 131072 --seconds 1 --steps 1 --warmup 3 --no-e2e --no-cpu --slab-parts 1`\n"
            "(one slab, so each launch covers all 131,072 streams of the bank).  Numbers under ncu are not bench values.\n\n")
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0]
        f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in want:
            if k in idx:
                f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
        dr = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) if "dram__bytes_read.sum" in idx else 0
        f.write("\nWarp stall reasons (average warps stalled per issue-active cycle), top 8:\n\n")
        vals = sorted(((float(r[idx[k]]), k) for k in stall if r[idx[k]] not in ("", "n/a")), reverse=True)[:8]
        for v, k in vals:
            f.write(f"* {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.3f}\n")
        f.write("\n")
print(open(os.path.join(out_dir, f"{rnd}_launches.md")).read()[:3000])
print(open(os.path.join(out_dir, f"{rnd}_ncu_summary.md")).read()[:5000])
