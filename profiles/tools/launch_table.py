"""Aggregates an `ncu --csv` launch list (several metrics per launch) into one row per kernel: launches, total / mean
time and the mean of every other metric. Usage: python profiles/tools/launch_table.py launches.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ix = {n: i for i, n in enumerate(H)}
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(H):
            continue
        k = r[ix["Kernel Name"]].split("(")[0][-48:]
        agg.setdefault(k, collections.OrderedDict()).setdefault(r[ix["Metric Name"]], []).append(
            float(r[ix["Metric Value"]].replace(",", "")))
    for k, m in agg.items():
        t = m.get("gpu__time_duration.sum", [0])
        print(f"{k}: launches {len(t)}, total {sum(t) / 1e6:.3f} ms, mean {sum(t) / len(t) / 1e6:.3f} ms")
        for n, v in m.items():
            if n != "gpu__time_duration.sum":
                print(f"    {n}: mean {sum(v) / len(v):.4g} sum {sum(v):.4g}")


if __name__ == "__main__":
    main(sys.argv[1])
