"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (and per grid shape).

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
"""
import collections
import csv
import re
import sys


def main(path, by_grid=False):
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines[start:]):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)[:80]
        key = (name, row["Grid Size"], row["Block Size"]) if by_grid else (name,)
        tot[key] += v
        cnt[key] += 1
    T = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {T / 1e3:.2f} ms of kernel time (ncu-serialised, cold cache)")
    print(f"# {'us_total':>10} {'share':>6} {'count':>6} {'us_avg':>8}  kernel" + (" grid block" if by_grid else ""))
    for k, v in tot.most_common(60):
        print(f"  {v:10.1f} {100 * v / T:5.1f}% {cnt[k]:6d} {v / cnt[k]:8.1f}  {' '.join(k)}")


if __name__ == "__main__":
    main(sys.argv[1], by_grid=len(sys.argv) > 2 and sys.argv[2] == "--by-grid")
