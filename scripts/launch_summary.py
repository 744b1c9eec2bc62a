#!/usr/bin/env python
"""Summarise ncu launch lists (--metrics gpu__time_duration.sum --csv) into a markdown table per file.
usage: launch_summary.py out.md title1=file1.csv [title2=file2.csv ...]"""
import collections
import csv
import re
import sys


def table(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("kj::", "")
        ns = float(r[col["Metric Value"]])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    setup = {"f32_to_bf16_kernel", "synth_rows_kernel"}
    enc = sum(v[1] for k, v in agg.items() if k not in setup)
    out = ["| kernel | launches | total us | us per launch | share (without one-time setup kernels) |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = "" if k in setup else f"{v[1] / enc:.3f}"
        out.append(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / 1e3 / v[0]:.1f} | {share} |")
    return out, tot


def main():
    out = sys.argv[1]
    lines = []
    for arg in sys.argv[2:]:
        title, path = arg.split("=", 1)
        t, tot = table(path)
        lines += [f"## {title}", ""] + t + [""]
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
