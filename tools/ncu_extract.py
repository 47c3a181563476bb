"""Turn one `ncu --set full` report into the small metric table committed under profiles/ and print where the warps wait.
usage: python tools/ncu_extract.py gpurun_out/prof_fe_umma_final.ncu-rep profiles/r02_ncu_fe_tcgen05_final_raw.csv
(how profiles/r02_ncu_fe_tcgen05*_raw.csv and the stall attribution in profiles/r02_fe_tcgen05.md were made)"""
import csv
import io
import subprocess
import sys

KEYS = """gpu__time_duration.sum smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
sm__warps_active.avg.per_cycle_active dram__bytes_read.sum dram__bytes_write.sum launch__registers_per_thread
launch__grid_size launch__block_size launch__shared_mem_per_block_dynamic""".split()
STALLS = ("long_scoreboard math_pipe_throttle not_selected wait short_scoreboard mio_throttle barrier sleeping "
          "branch_resolving").split()


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    keys = KEYS + [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in STALLS]
    with open(dst, "w") as f:
        f.write("metric,value,unit\n")
        for k in keys:
            if k in d:
                f.write(f"{k},{d[k][0]},{d[k][1]}\n")
    rows = page(rep, "source")
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]

    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    print("stall samples:", int(sum(num(r[ci["# Samples"]]) for r in data)))
    for col in ("stall_long_sb", "stall_short_sb", "stall_math", "stall_wait", "stall_mio", "stall_barrier"):
        print(f"  {col}: {int(sum(num(r[ci[col]]) for r in data))}")
    for col in ("stall_long_sb", "stall_short_sb", "# Samples"):
        print(f"top {col}:")
        for r in sorted(data, key=lambda r: -num(r[ci[col]]))[:10]:
            print(f"  {int(num(r[ci[col]])):6d}  {r[ci['Source']][:90]}")


if __name__ == "__main__":
    main()
