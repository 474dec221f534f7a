#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/ncu_summarize.py gpurun_out/launches.csv > profiles/r01_launches_summary.md"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void |^\(anonymous namespace\)::|<unnamed>::", "", name)
        rows.append((name, ns))
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"# launch list summary: {len(rows)} launches, {total / 1e6:.2f} ms total (ncu-serialised, cold-cache: compare SHARES)\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n[:90]}` | {c} | {t / 1e6:.3f} | {100 * t / total:.1f}% | {t / c / 1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
