"""List every implicit host<->device synchronisation inside one pre-training step (torch sync-debug mode)."""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from xfm_b200.accelerator import B200DDPAccelerator, FlatAdamW  # noqa: E402
from xfm_b200.model_pretrain import XFM  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model = XFM(bench.base_config(), init=bench.gpu_init(dev, 0), device=dev).train()
opt = FlatAdamW(model, lr=1e-4, weight_decay=0.01, lr_mult=2.0)
acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0))
b = {k: v.to(dev) for k, v in bench.make_host_batch(16, 40, 15, model.cfg["vocab_size"], 224, 100).items()}


def step():
    out = model(b["image"], b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
                masked_ids=b["masked_ids"], ret_mim_loss=True, data_source="image")
    loss = out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + out["loss_mim"]
    acc.backward_step(loss, opt)
    acc.optimizer_step(opt, model)
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    step()
torch.cuda.set_sync_debug_mode("default")
import traceback
seen = {}
for x in w:
    key = (x.filename, x.lineno)
    seen[key] = seen.get(key, 0) + 1
for (f, l), n in seen.items():
    print(n, f, l)
print("total sync warnings", len(w))
