#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).

    python scripts/summarize_launches.py profiles/r01_launches.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
    tot = OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        ns = float(r[14].replace(",", ""))
        grid = r[8]
        k = tot.setdefault(name, {"n": 0, "ns": 0.0, "grids": set()})
        k["n"] += 1
        k["ns"] += ns
        k["grids"].add(grid)
    total = sum(k["ns"] for k in tot.values())
    print("| kernel | launches | total us | share | grids |")
    print("|---|---|---|---|---|")
    for name, k in tot.items():
        grids = sorted(k["grids"])
        g = ", ".join(grids) if len(grids) <= 3 else f"{len(grids)} shapes"
        print(f"| `{name}` | {k['n']} | {k['ns'] / 1e3:.1f} | {100 * k['ns'] / total:.2f} % | {g} |")
    print(f"\ntotal {total / 1e3:.1f} us over {len(rows)} launches")


if __name__ == "__main__":
    main()
