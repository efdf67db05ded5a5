#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share.

    python scripts/summarize_launches.py launches.csv [--between KERNEL]   # last segment between two KERNEL launches
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    between = sys.argv[sys.argv.index("--between") + 1] if "--between" in sys.argv else None
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")))
            for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    if between:
        idx = [i for i, (n, _) in enumerate(rows) if between in n]
        if len(idx) >= 2:
            rows = rows[idx[-2] + 1: idx[-1] + 1]
    tot = sum(t for _, t in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in rows:
        k = re.sub(r"\(.*", "", re.sub(r"^void ", "", n)).replace("fact::", "")
        agg[k][0] += 1
        agg[k][1] += t
    print(f"launches {len(rows)}  total {tot / 1e6:.3f} ms")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e6:8.3f} ms {100 * t / tot:5.1f}%  n={c:4d}  avg={t / c / 1e3:8.1f} us  {k[:100]}")


if __name__ == "__main__":
    main()
