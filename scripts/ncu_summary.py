#!/usr/bin/env python
"""Summarise .ncu-rep files (read here, no GPU needed) into a small CSV/markdown for profiles/.
usage: ncu_summary.py out.md rep1.ncu-rep [rep2 ...]"""
import csv
import io
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
]


def main():
    out = sys.argv[1]
    lines = ["| report | kernel | " + " | ".join(k for _, k in KEEP) + " |", "|---|---|" + "---|" * len(KEEP)]
    for rep in sys.argv[2:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
            vals = []
            for m, _ in KEEP:
                i = col.get(m)
                vals.append("" if i is None else f"{r[i]} {units[i]}".strip())
            lines.append(f"| {rep.split('/')[-1]} | `{name}` | " + " | ".join(vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
