"""Per-kernel SASS opcode histogram of xfm_b200/libxfm_b200.so (cuobjdump -sass): proof, without the binary, of which
kernels issue tcgen05 MMAs (UTCHMMA), TMEM loads / stores (LDTM / STTM), TMA loads / stores (UTMALDG / UTMASTG), cluster
barriers and legacy mma.sync (HMMA).   python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "xfm_b200", "libxfm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "MUFU.EX2", "RED", "ATOM", "LDGSTS")
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEY:
            if op.startswith(k):
                per[cur][k] += 1
                if k in ("UTCHMMA", "UTCBAR", "UTMALDG") and ".2CTA" in op:
                    per[cur][k + ".2CTA"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print(f"# {os.path.relpath(so, ROOT)}: {len(per)} kernels (cuobjdump -sass, sm_100a)")
print("# instruction counts per kernel; only kernels with at least one tensor-core / TMEM / TMA / mma.sync instruction are listed")
for (name, c), dn in zip(per.items(), demangle):
    for k, v in c.items():
        tot[k] += v
    keys = [k for k in c if k != "_total" and k not in ("SYNCS", "RED", "ATOM", "MUFU.EX2")]
    if not keys:
        continue
    short = re.sub(r"\(.*", "", dn).replace("void ", "")[:90]
    print(f"{short:92s} total={c['_total']:6d}  " + "  ".join(f"{k}={c[k]}" for k in sorted(c) if k != "_total"))
print("# whole library: " + "  ".join(f"{k}={tot[k]}" for k in sorted(tot)))
