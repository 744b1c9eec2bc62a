#!/usr/bin/env python
"""Top sampled SASS instructions of one kernel in an .ncu-rep (source page). usage: ncu_hot.py rep kernel_regex [idx] [topn]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[which]
e = starts[which + 1] if which + 1 < len(starts) else len(rows)
print(rows[s][1][:140])
hdr = rows[s + 1]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
stall_cols = [h for h in hdr if h.startswith("stall_")]
ns = "# Samples"
tot = sum(int(r[col[ns]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
if stall_cols:
    agg = {h: sum(int(r[col[h]] or 0) for r in body) for h in stall_cols}
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
for i, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][col[ns]] or 0))[:topn]:
    st = {h[6:]: int(r[col[h]] or 0) for h in stall_cols if int(r[col[h]] or 0) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f'{r[col[ns]]:>7} #{i:<5} {r[col["Source"]].strip()[:80]:80s} exec={r[col["Instructions Executed"]]:>8} {top}')
