"""dQ.q / dK.k balance of the tcgen05 ViT backward with the forward results of either kernel family (diagnostic)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfm_b200 import lib
from xfm_b200.encoders import closed_form_rel_index
B, H, N = 96, 12, 197
D = H * 64
M = B * N
g = torch.Generator(device="cuda").manual_seed(2)
qkv = (torch.randn(M, 3 * D, device="cuda", generator=g) * 0.6).bfloat16()
table = torch.randn(732, H, device="cuda", generator=g)
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
idx = closed_form_rel_index(14).cuda()
bias = table[idx.view(-1)].view(N, N, H).permute(2, 0, 1).contiguous()
ld = (N + 7) // 8 * 8
bias_p = torch.zeros(H, N, ld, device="cuda")
bias_p[:, :, :N] = bias
out, lse = lib.attention_fwd(q, k, v, B, H, N, N, 0.125, rel_table=table, rel_window=14)
out2, lse2 = lib.attention_fwd(q, k, v, B, H, N, N, 0.125, bias=bias_p, allow_tc=False)
f = qkv[:4 * N].float().view(4, N, 3, H, 64).permute(2, 0, 3, 1, 4)
s = (f[0] * 0.125) @ f[1].transpose(-1, -2) + bias
ref = torch.logsumexp(s, -1)
print("lse tc - ref: max", float((lse[:4] - ref).abs().max()), "mean", float((lse[:4] - ref).mean()))
print("lse mma - ref: max", float((lse2[:4] - ref).abs().max()), "mean", float((lse2[:4] - ref).mean()))
oref = (torch.softmax(s, -1) @ f[2]).permute(0, 2, 1, 3).reshape(4 * N, D)
print("out tc - ref max", float((out[:4 * N].float() - oref).abs().max()), "out mma - ref", float((out2[:4 * N].float() - oref).abs().max()))
dout = (torch.randn(M, D, device="cuda", generator=g)).bfloat16()
for name, (o, l) in (("tc fwd", (out, lse)), ("mma fwd", (out2, lse2))):
    for tc in (True, False):
        dqkv = torch.empty_like(qkv)
        kw = dict(rel_table=table, rel_window=14) if tc else dict(bias=bias_p, allow_tc=False)
        lib.attention_bwd(dout, q, k, v, o, l, B, H, N, N, 0.125, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], **kw)
        a = float((dqkv[:, :D].float() * q.float()).sum())
        b_ = float((dqkv[:, D:2 * D].float() * k.float()).sum())
        print(name, "tc bwd" if tc else "mma bwd", "a", a, "b", b_, "diff", a - b_)
