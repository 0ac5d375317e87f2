"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total/avg time, share."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        name = r["Kernel Name"]
        name = re.sub(r"\(.*$", "", name)
        rows.append((name, val * scale))
    agg = OrderedDict()
    for n, t in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    lines = ["%-90s %7s %12s %10s %7s" % ("kernel", "launch", "total_us", "avg_us", "share")]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-90s %7d %12.1f %10.2f %6.1f%%" % (n[:90], c, t, t / c, 100 * t / tot))
    lines.append("%-90s %7d %12.1f" % ("TOTAL", len(rows), tot))
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
