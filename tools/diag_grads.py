"""Diagnostic: per-parameter gradient error of the CUDA path vs the reference fixtures (run on the GPU box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import xfm_oracle as O
from xfm_b200.model_pretrain import XFM

name = sys.argv[1] if len(sys.argv) > 1 else "tiny_vq.pt"
g = torch.load(os.path.join(ROOT, "tests/golden", name), weights_only=False)
cfg = dict(g["cfg"])
model = XFM(cfg, init=lambda n, s: O.make_tensor(n, s, 0), device="cuda").eval()
b = {k: v.cuda() for k, v in O.make_batch(cfg, g["B"], L=g["L"], M=g["M"], seed=1, image_uniform=g["image_uniform"]).items()}
model._forced_negatives = (g["image_neg_idx"], g["text_neg_idx"])
model._forced_masks = g["ids_mask"]
out = model(b["image"], b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
            masked_ids=b["masked_ids"], ret_mim_loss=True, data_source="image")
for k, v in g["losses"].items():
    print(f"{k}: mine {float(out[k]):.6f} ref {v:.6f} rel {abs(float(out[k]) - v) / max(1, abs(v)):.2e}")
(out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + out["loss_mim"]).backward()
params = dict(model.named_parameters())
for n, ref in g.get("grads", {}).items():
    mine = params[n].grad
    if mine is None:
        print("NONE", n); continue
    d = (mine.float().cpu() - ref).abs()
    i = int(d.argmax())
    print(f"{float(d.max()) / max(float(ref.abs().max()), 1e-8):.3e}  refmax {float(ref.abs().max()):.3e}  at {i} "
          f"(mine {float(mine.flatten()[i]):.4e} ref {float(ref.flatten()[i]):.4e})  {n} {tuple(ref.shape)}")
