"""Summarise an .ncu-rep: headline metrics, warp-state breakdown, and the hottest SASS instructions.
Usage: python tools/ncu_summary.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__icc_request_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=====")
    for k in KEYS:
        if k in d: print(f"{k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
    print("-- warp states (avg warps per issue-cycle... pct of warp-active cycles)")
    st = []
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            st.append((float(d[k] or 0), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    for v, k in sorted(st, reverse=True)[:12]: print(f"   {k:30s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# find header row
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c for c in r))
h = rows[hi]
ci = {c: i for i, c in enumerate(h)}
samp = next(c for c in h if c.startswith("Warp Stall Sampling (All"))
ex = next((c for c in h if c.startswith("Instructions Executed")), None)
data = rows[hi + 1:]
tot = sum(float(r[ci[samp]] or 0) for r in data if len(r) > ci[samp])
print("total samples", tot, "rows", len(data))
best = sorted(range(len(data)), key=lambda i: -float(data[i][ci[samp]] or 0) if len(data[i]) > ci[samp] else 0)[:topn]
for i in sorted(best):
    r = data[i]
    print(f"{i:6d} {float(r[ci[samp]])/tot*100:6.2f}%  ex={r[ci[ex]] if ex else ''}  {r[ci['Source']][:110]}")
open("/tmp/ncu_source.csv", "w").write(src)
