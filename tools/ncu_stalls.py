"""Per-instruction stall summary from an `ncu --set full --import-source on` report (run here, no GPU needed):
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [top_n]
Prints the stall-reason totals and the SASS instructions that collected the most warp-stall samples."""
import csv
import subprocess
import sys


def main(path, top_n=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        h = b["hdr"]
        iS, isrc, iex = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
        cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[iS]) for r in b["rows"]) or 1
        nex = sum(int(r[iex] or 0) for r in b["rows"])
        print(f"== {b['name'][:70]}: {tot} samples, {len(b['rows'])} SASS instructions, {nex} warp-instructions executed")
        agg = {}
        for r in b["rows"]:
            for i in cols:
                agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
        for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
            print(f"   {k:26s} {v:7d} {100 * v / tot:5.1f}%")
        top = sorted(enumerate(b["rows"]), key=lambda x: -int(x[1][iS]))[:top_n]
        for idx, r in sorted(top):
            reasons = sorted(((int(r[i] or 0), h[i][6:]) for i in cols), reverse=True)[:2]
            print(f"{idx:5d} {int(r[iS]):6d} {100 * int(r[iS]) / tot:5.1f}% ex={r[iex]:>8s} {r[isrc].strip()[:64]:64s} {reasons}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
