"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares, not absolutes)."""
import collections
import csv
import re
import sys


def main(path, skip=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for i, row in enumerate(csv.DictReader(lines)):
        if i < skip:
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v * 1e6 if unit in ("s", "second") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"{'kernel':58s} {'launches':>8s} {'total_us':>11s} {'avg_us':>9s} {'share':>6s}")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:58]:58s} {n:8d} {t:11.1f} {t / n:9.1f} {t / tot:6.3f}")
    print(f"{'TOTAL':58s} {'':8s} {tot:11.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
