"""A few launches of the tcgen05 ViT attention forward at the pre-training shape (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfm_b200 import lib as L  # noqa: E402

B, H, N = 96, 12, 197
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * N, 3 * D, device="cuda", generator=g).bfloat16()
table = torch.randn(732, H, device="cuda", generator=g)
out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    L.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, N, N, 0.125, rel_table=table, rel_window=14, out=out)
torch.cuda.synchronize()
print("ok")
