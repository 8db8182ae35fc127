#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.

    python profiles/summarize_launches.py gpurun_out/launches.csv [--steps K] [--warmup W] > profiles/rNN_launches.md

bench.py runs W + W + K + K (priming) + K (timed, HBM-resident) + K (timed, host buffers) guided steps; every
step ends with exactly one `guided_step_vec4` launch, which is used to cut the list into steps.  The
summary covers the first timed region (the `value` leg).  ncu times are cold-cache and serialised:
the SHARES are what must agree with the live CUDA-event numbers, not the absolutes.
"""
import argparse
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("b2e::", "").replace("void ", "")
    return name[:90]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--steps", type=int, default=2, help="guided denoising steps in the summarised region")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--before", type=int, default=None,
                    help="guided denoising steps before the timed region (round-2 bench.py: (warmup + 1) passes x --denoise-steps)")
    a = ap.parse_args()
    rows = []
    with open(a.csv, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((int(r["ID"]), r["Kernel Name"], ns))
    # cut into steps at guided_step launches
    ends = [i for i, r in enumerate(rows) if "guided_step" in r[1]]
    K, W = a.steps, a.warmup
    n_before = a.before if a.before is not None else W + min(W, K) + K + K   # warm-up + priming steps before the timed region
    if len(ends) < n_before + K:
        print(f"only {len(ends)} guided_step launches in the list (need {n_before + K}); summarising everything", file=sys.stderr)
        lo, hi = 0, len(rows)
    else:
        lo = ends[n_before - 1] + 1
        hi = ends[n_before + K - 1] + 1
    region = rows[lo:hi]
    tot = sum(r[2] for r in region)
    by = defaultdict(lambda: [0, 0.0])
    for _, name, ns in region:
        by[short(name)][0] += 1
        by[short(name)][1] += ns
    print(f"# ncu launch list: {K} guided denoising steps of the timed region of bench.py (launch IDs {region[0][0]}..{region[-1][0]})\n")
    print(f"{len(region)} launches, {tot / 1e6:.3f} ms serialised device time ({tot / 1e6 / max(K, 1):.3f} ms/step), "
          f"{len(region) / max(K, 1):.0f} launches/step\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, ns) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f} % | {ns / n / 1e3:.1f} |")
    print("\nTop 25 single launches:\n")
    print("| id | kernel | us |")
    print("|---:|---|---:|")
    for i, name, ns in sorted(region, key=lambda r: -r[2])[:25]:
        print(f"| {i} | `{short(name)}` | {ns / 1e3:.1f} |")


if __name__ == "__main__":
    main()
