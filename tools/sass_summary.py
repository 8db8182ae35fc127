#!/usr/bin/env python
"""Per-kernel SASS evidence that the contraction kernels are Blackwell-native: counts of the tcgen05 / TMA / TMEM mnemonics
in every sm_100a cubin function of libb200edit.so (cuobjdump -sass).  Regenerates profiles/sass_summary.txt.

    python tools/sass_summary.py [out.txt]

UTCHMMA(.2CTA) = tcgen05.mma (kind::f16, one / two CTAs), UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld
(TMEM -> registers), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, UBLKCP = cp.async.bulk (here: the shared::cta ->
shared::cluster copy of the cluster split-K partial tiles), UCGABAR = barrier.cluster, HMMA = legacy warp-level mma.sync (must be 0)."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "diffusion-image-editing_b200", "b200edit", "libb200edit.so")
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "MUFU", "LDG", "STG", "UBLKCP", "UCGABAR_ARV"]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "sass_summary.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
    rows, cur, arch = [], None, None
    counts = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                rows.append((cur, counts))
            cur, counts = m.group(1), collections.Counter()
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op, suffix = m.group(1), m.group(2)
            for k in KEYS:
                if op == k:
                    counts[k] += 1
                    if k in ("UTCHMMA", "UTMALDG") and ".2CTA" in suffix:
                        counts[k + ".2CTA"] += 1
    if cur:
        rows.append((cur, counts))
    total = collections.Counter()
    lines = [f"# SASS summary of {os.path.relpath(LIB, REPO)} ({arch}); regenerate with tools/sass_summary.py", "#",
             "# " + " ".join(f"{k:>8s}" for k in ["UTCHMMA", "(.2CTA)", "UTMALDG", "(.2CTA)", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "UBLKCP", "UCGABAR", "HMMA", "MUFU"]) + "  kernel"]
    for name, c in sorted(rows, key=lambda r: -r[1]["UTCHMMA"]):
        total.update(c)
        if not (c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"] or c["LDTM"]):
            continue
        d = demangle(name)
        d = re.sub(r"\(.*", "", d)
        lines.append("  " + " ".join(f"{c[k]:8d}" for k in ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "UBLKCP", "UCGABAR_ARV", "HMMA", "MUFU"]) + "  " + d)
    lines.append("#")
    lines.append(f"# all {len(rows)} functions: " + ", ".join(f"{k} {total[k]}" for k in ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UBLKCP", "HMMA"]))
    lines.append("# HMMA (warp-level mma.sync) must be 0: every dense contraction goes through tcgen05.mma with the accumulator in TMEM")
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[-3:]))
    print("wrote", out_path)


if __name__ == "__main__":
    main()
