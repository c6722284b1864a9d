"""Key metrics per launch from an `ncu --set full` report (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_summary.txt
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {path}: {len(data)} profiled launches (ncu --set full --clock-control none)")
    for r in data:
        name = r[hdr.index("Kernel Name")]
        print(f"\n{name[:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"    {k:70s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
