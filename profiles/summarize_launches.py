"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
        agg[name][0] += 1
        agg[name][1] += v
    ours = {k: v for k, v in agg.items() if "lcc::" in k}
    other = {k: v for k, v in agg.items() if "lcc::" not in k}
    tot = sum(v[1] for v in ours.values())
    print(f"total device time in lcc:: kernels: {tot / 1e3:.2f} ms over {sum(v[0] for v in ours.values())} launches "
          f"(other kernels, i.e. torch synthetic-weight generation and copies: {sum(v[1] for v in other.values()) / 1e3:.2f} ms)\n")
    print("| share | launches | avg us | kernel |\n|---:|---:|---:|---|")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v[1] / tot * 100:.2f}% | {v[0]} | {v[1] / v[0]:.2f} | `{k.strip()}` |")


if __name__ == "__main__":
    main(sys.argv[1])
