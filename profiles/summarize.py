"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel: count, mean, share."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    acc = defaultdict(list)
    for r in rows:
        name = re.sub(r"\(.*", "", r[k]).replace("void ", "")
        acc[name].append(float(r[v].replace(",", "")))
    total = sum(sum(x) for x in acc.values())
    print(f"{'kernel':28s} {'launches':>8s} {'mean us':>10s} {'sum us':>10s} {'share':>7s}")
    for name, x in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        print(f"{name:28s} {len(x):8d} {sum(x) / len(x) / 1e3:10.1f} {sum(x) / 1e3:10.1f} {sum(x) / total:7.1%}")
    print(f"{'total':28s} {sum(len(x) for x in acc.values()):8d} {'':10s} {total / 1e3:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
