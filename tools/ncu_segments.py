"""Aggregate ncu source-page samples of a kernel into segments delimited by BAR.SYNC (per-segment stall mix).
Usage: python tools/ncu_segments.py /tmp/ncu_source.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c for c in r))
h = rows[hi]; ci = {c: i for i, c in enumerate(h)}
data = rows[hi + 1:]
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(float(r[ci['Warp Stall Sampling (All Samples)']] or 0) for r in data)
seg_start = 0
def flush(a, b):
    n = sum(float(r[ci['Warp Stall Sampling (All Samples)']] or 0) for r in data[a:b])
    if n / tot < 0.004: return
    ex = max(int(float(r[ci['Instructions Executed']] or 0)) for r in data[a:b])
    exsum = sum(int(float(r[ci['Instructions Executed']] or 0)) for r in data[a:b])
    mix = collections.Counter()
    for r in data[a:b]:
        for s in stalls: mix[s[6:]] += float(r[ci[s]] or 0)
    top = ", ".join(f"{k}={v/n*100:.0f}%" for k, v in mix.most_common(6))
    print(f"[{a:5d},{b:5d}) n_instr={b-a:5d} ex_max={ex:9d} inst_exec={exsum:12d} samples={n/tot*100:5.1f}%  {top}")
for i, r in enumerate(data):
    if 'BAR.SYNC' in r[ci['Source']]:
        flush(seg_start, i + 1); seg_start = i + 1
flush(seg_start, len(data))
