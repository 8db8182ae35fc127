#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report as a markdown table:  summarize_ncu.py <file.ncu-rep | raw.csv>"""
import csv, subprocess, sys

WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("sm__cycles_active.avg", "SM active cycles")]
if sys.argv[1].endswith(".csv"):      # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`)
    out = open(sys.argv[1]).read()
else:
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print("| id | kernel | " + " | ".join(f"{n} [{units[idx[m]]}]" if units[idx[m]] else n for m, n in WANT if m in idx) + " |")
print("|---|---|" + "---:|" * sum(1 for m, _ in WANT if m in idx))
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:48]
    vals = []
    for m, _ in WANT:
        if m in idx:
            v = r[idx[m]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            vals.append(v)
    print(f"| {r[idx['ID']]} | `{name}` | " + " | ".join(vals) + " |")
