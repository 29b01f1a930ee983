#!/usr/bin/env python
"""Top stalled SASS instructions of an ncu report: python scripts/ncu_hot.py report.ncu-rep [N]."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
data = rows[rows.index(hdr) + 1:]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
num = lambda s: int(s) if s.strip().isdigit() else 0
data = [r for r in data if len(r) > i_samp]
tot = sum(num(r[i_samp]) for r in data)
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + num(r[i])
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -num(r[i_samp]))[:n]:
    st = {hdr[i]: num(r[i]) for i in stall_cols if num(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{num(r[i_samp]):6d} {100 * num(r[i_samp]) / max(tot, 1):5.1f}% exec={r[i_exec]:>8s} {r[i_src][:72]:72s} {st}")
