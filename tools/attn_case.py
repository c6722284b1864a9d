"""A few launches of the tcgen05 ViT attention kernels at the pre-training shape (for ncu).  usage: attn_case.py [fwd|bwd] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfm_b200 import lib as L  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, H, N = 96, 12, 197
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * N, 3 * D, device="cuda", generator=g).bfloat16()
table = torch.randn(732, H, device="cuda", generator=g)
out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
dout = torch.randn(B * N, D, device="cuda", generator=g).bfloat16()
dqkv = torch.empty_like(qkv)
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
o, lse = L.attention_fwd(q, k, v, B, H, N, N, 0.125, rel_table=table, rel_window=14, out=out)
for _ in range(reps):
    if what == "fwd":
        L.attention_fwd(q, k, v, B, H, N, N, 0.125, rel_table=table, rel_window=14, out=out)
    else:
        L.attention_bwd(dout, q, k, v, out, lse, B, H, N, N, 0.125, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                        rel_table=table, rel_window=14)
torch.cuda.synchronize()
print("ok")
