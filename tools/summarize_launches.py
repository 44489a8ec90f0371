#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (time, count, share).

    python tools/summarize_launches.py gpurun_out/launches.csv "title" > profiles/rNN_launches.md
"""
import collections
import csv
import re
import sys


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        agg[name[:90]][0] += 1
        agg[name[:90]][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# {title}\n")
    print(f"source: `{path}` — {n} launches, {tot / 1e3:.2f} ms of kernel time (ncu per-launch times are cold-cache and "
          "serialised: shares are meaningful, absolutes are not)\n")
    print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.2f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main()
