"""profiles/gemm_traffic.json (read by bench.py's roofline.traffic) from an `ncu --set full` capture of the GEMM kernels
inside bench.py (run here, no GPU needed):
    python tools/gemm_traffic.py gpurun_out/r01f_bench_gemm_pair.ncu-rep "<the ncu command line>" > profiles/gemm_traffic.json
"""
import csv
import json
import subprocess
import sys


def main(path, source):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def val(r, key):
        i = hdr.index(key)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    per = []
    for r in data:
        per.append(dict(kernel=r[hdr.index("Kernel Name")][:60], us=round(val(r, "gpu__time_duration.sum") /
                        (1e3 if units[hdr.index("gpu__time_duration.sum")] == "ns" else 1), 3),
                        dram_read=val(r, "dram__bytes_read.sum"), dram_write=val(r, "dram__bytes_write.sum"),
                        tensor_pct=val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")))
    n = max(len(per), 1)
    print(json.dumps(dict(source=source, kernel="gemm_tcgen05_pair_kernel", launches=len(per),
                          dram_bytes_per_launch=sum(p["dram_read"] + p["dram_write"] for p in per) / n, per_launch=per), indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
