"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals + sequence."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000.0 if u == "ns" else (v * 1000.0 if u == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        seq.append((name[:90], v, row.get("Grid Size"), row.get("Block Size")))
    return seq


def main():
    seq = load(sys.argv[1])
    agg = collections.OrderedDict()
    for n, v, _, _ in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v, _, _ in seq)
    mine = sum(t for n, (c, t) in agg.items() if n.startswith("mvb::"))
    print(f"# {len(seq)} launches, {tot:.1f} us total (serialised, cold-cache), libmvb kernels {mine:.1f} us = {100*mine/tot:.1f} %")
    print(f"# {'us':>9} {'share':>6} {'n':>4}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:11.1f} {100*t/tot:5.1f}% {c:4d}  {n}")
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            for i, (n, v, g, b) in enumerate(seq):
                f.write(f"{i:4d} {v:8.1f} {g} {b} {n}\n")


if __name__ == "__main__":
    main()
