"""Aggregates an `ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --csv` launch list by kernel: executed warp
instructions are the work measure that survives overlapping streams (time under ncu is serialised and cold-cache)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(name, {"n": 0, "inst": 0.0, "us": 0.0})
        if row["Metric Name"].startswith("smsp__inst_executed"):
            a["inst"] += v
            a["n"] += 1
        else:
            unit = row["Metric Unit"]
            a["us"] += v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    tot_i = sum(a["inst"] for a in agg.values())
    tot_t = sum(a["us"] for a in agg.values())
    print(f"{'kernel':44s} {'launches':>8s} {'Minst':>10s} {'inst%':>6s} {'us':>10s} {'time%':>6s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["inst"]):
        print(f"{k[:44]:44s} {a['n']:8d} {a['inst'] / 1e6:10.1f} {a['inst'] / tot_i:6.3f} {a['us']:10.1f} {a['us'] / tot_t:6.3f}")
    print(f"{'TOTAL':44s} {'':8s} {tot_i / 1e6:10.1f} {'':6s} {tot_t:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
