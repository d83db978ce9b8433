#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of device time)."""
import collections
import csv
import sys


def summarize(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(",", "")) / 1e3  # ns -> us
    total = sum(v[1] for v in agg.values())
    lines = ["| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / total))
    return "\n".join(lines)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
