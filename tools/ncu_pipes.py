"""Every "% of peak" pipe / memory-unit utilisation metric of the launches in an `ncu --set full` report, sorted — the
one-screen answer to "which unit is the kernel bound by" (run here or on the GPU box):
    python tools/ncu_pipes.py x.ncu-rep [min_pct]
"""
import csv
import subprocess
import sys


def main(path, min_pct=5.0):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        print(f"== {r[hdr.index('Kernel Name')][:100]}  ({r[hdr.index('gpu__time_duration.sum')]} {units[hdr.index('gpu__time_duration.sum')]})")
        vals = []
        for i, (h, u) in enumerate(zip(hdr, units)):
            if u == "%" and ("pct_of_peak" in h):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if v >= min_pct:
                    vals.append((v, h))
        for v, h in sorted(vals, reverse=True):
            print(f"   {v:7.2f} %  {h}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 5.0)
