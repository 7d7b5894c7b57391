#!/usr/bin/env python
"""ncu launch list CSV (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) -> markdown table with
each kernel's share:  python profiles/launch_list.py gpurun_out/c35_launches.csv > profiles/rX_launches_step.md"""
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("b200ps_impl::", "").replace("<unnamed>::", "")
    return name.split("(")[0][:70]


def main():
    rows = []
    with open(sys.argv[1], newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            ns = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit") in ("us", "usecond"):
                ns *= 1e3
            rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], ns / 1e3))
    total = sum(r[3] for r in rows)
    print("| # | kernel | grid | block | time (us) | share |")
    print("|---|---|---|---|---|---|")
    for i, (n, g, b, us) in enumerate(rows):
        print("| %d | `%s` | %s | %s | %.1f | %.1f%% |" % (i, n, g, b, us, 100 * us / total))
    print("\ntotal %.1f us over %d launches" % (total, len(rows)))


if __name__ == "__main__":
    main()
