"""Kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`): markdown table on stdout.

    python tools/launch_shares.py profiles/r1_launches.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*$", "", r[k]).replace("void ", "").replace("<unnamed>::", "").strip()
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + float(r[v].replace(",", "")) / 1e3)
    total = sum(t for _, t in agg.values())
    print(f"{sum(n for n, _ in agg.values())} launches, {total:.0f} us in total\n")
    print("| kernel | launches | total µs | share |\n|---|---|---|---|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {t:.0f} | {100 * t / total:.1f} % |")


if __name__ == "__main__":
    main()
