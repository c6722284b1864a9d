"""One profiled pre-training step (same workload as bench.py) between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from xfm_b200.accelerator import B200DDPAccelerator, FlatAdamW  # noqa: E402
from xfm_b200.model_pretrain import XFM  # noqa: E402

B = int(os.environ.get("XFM_BENCH_PAIRS", "0"))
warm = int(os.environ.get("XFM_PROFILE_WARMUP", "2"))
task = os.environ.get("XFM_PROFILE_CONFIG", "pretrain")      # pretrain | retrieval | nlvr | vqa
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
if task == "pretrain":
    model = XFM(bench.base_config(), init=bench.gpu_init(dev, 0), device=dev).train()
    opt = FlatAdamW(model, lr=1e-4, weight_decay=0.01, lr_mult=2.0)
    acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0))
    b = {k: v.to(dev) for k, v in bench.make_host_batch(B or 96, 40, 15, model.cfg["vocab_size"], 224, 100).items()}

    def loss_of():
        out = model(b["image"], b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
                    masked_ids=b["masked_ids"], ret_mim_loss=True, data_source="image")
        return out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + out["loss_mim"]
else:
    wl = bench.finetune_workload(task, int(os.environ.get("XFM_PROFILE_RES", "384")), B, dev)
    _, model, opt, acc = wl["build"]()
    b = {k: v.to(dev) for k, v in wl["host"][0].items()}

    def loss_of():
        return wl["loss_fn"](model, b)


def step():
    loss = loss_of()
    acc.backward_step(loss, opt)
    acc.optimizer_step(opt, model)
    return loss


for _ in range(warm):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled step loss", float(loss.detach()))
