"""profiles/gemm_traffic.json (read by bench.py's roofline.traffic): DRAM bytes of EVERY tcgen05 GEMM launch of one
pre-training step, from an ncu pass over tools/profile_step.py:

    ncu --profile-from-start off --clock-control none -k regex:gemm_tcgen05 \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv \
        --log-file gpurun_out/gemm_traffic.csv python tools/profile_step.py
    python tools/gemm_traffic.py gpurun_out/gemm_traffic.csv > profiles/gemm_traffic.json      (run here, no GPU needed)

(the same counters an `--set full` capture reports; a single-pass metric list keeps the 800 launches of a step affordable).
Round 1 averaged the 12 smallest launches of a step; this is the mean over all of them."""
import collections
import csv
import json
import re
import sys


def main(path):
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    per = collections.defaultdict(dict)
    names = {}
    for row in csv.DictReader(lines[start:]):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"].lower()
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        per[row["ID"]][row["Metric Name"]] = v
        names[row["ID"]] = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    rows = [dict(kernel=names[i], us=m.get("gpu__time_duration.sum", 0.0), dram_read=m.get("dram__bytes_read.sum", 0.0),
                 dram_write=m.get("dram__bytes_write.sum", 0.0)) for i, m in per.items()]
    n = max(len(rows), 1)
    tot = sum(r["dram_read"] + r["dram_write"] for r in rows)
    by = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows:
        b = by[r["kernel"]]
        b[0] += 1
        b[1] += r["dram_read"] + r["dram_write"]
        b[2] += r["us"]
    big = sorted(rows, key=lambda r: -r["us"])[:12]
    print(json.dumps(dict(
        source="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
               "-k regex:gemm_tcgen05 --profile-from-start off python tools/profile_step.py (one pre-training step, B = 96)",
        kernel="gemm_tcgen05_pair_kernel / gemm_tcgen05_kernel", launches=len(rows), dram_bytes_per_launch=tot / n,
        dram_bytes_per_step=tot, us_per_step_serialised=sum(r["us"] for r in rows),
        per_kernel={k: dict(launches=v[0], dram_bytes_per_launch=v[1] / v[0], us_per_launch=v[2] / v[0]) for k, v in by.items()},
        longest_launches=big), indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
