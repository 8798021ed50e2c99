"""Per-region instruction and stall-sample histogram from `ncu --page source --csv`.
Usage: python scripts/ncu_source_hist.py report.ncu-rep [nbins]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(r[isrc].strip(), int(r[iex]), int(r[ismp])) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
tot_i, tot_s = sum(d[1] for d in data), sum(d[2] for d in data)
print("instructions", tot_i, "samples", tot_s, "sass lines", len(data))
# segment by execution-count plateaus
seg_start, prev = 0, data[0][1]
def flush(a, b):
    n = sum(d[1] for d in data[a:b]); s = sum(d[2] for d in data[a:b])
    ops = {}
    for d in data[a:b]:
        op = d[0].split()[0] if not d[0].startswith("@") else d[0].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:8]
    print(f"lines {a:5d}-{b:5d} n={b-a:4d} exec/line~{data[a][1]:>10d} inst {100*n/tot_i:5.1f}% samples {100*s/tot_s:5.1f}%  {top}")
for i, d in enumerate(data):
    if abs(d[1] - prev) > 0.2 * max(prev, 1) and i - seg_start >= 8:
        flush(seg_start, i); seg_start = i
    prev = d[1]
flush(seg_start, len(data))
print("--- top stall lines")
for d in sorted(data, key=lambda x: -x[2])[:25]:
    print(f"{d[2]:6d} {100*d[2]/tot_s:5.1f}%  exec {d[1]:>9d}  {d[0][:100]}")
