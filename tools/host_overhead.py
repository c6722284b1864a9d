"""How long the host takes to ISSUE one pre-training step (no sync) vs how long the GPU takes to run it.
If issue time ~ step time the step is launch-bound and kernel speed-ups will not show."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from xfm_b200 import lib as L  # noqa: E402
from xfm_b200.accelerator import B200DDPAccelerator, FlatAdamW  # noqa: E402
from xfm_b200.model_pretrain import XFM  # noqa: E402

B = int(os.environ.get("XFM_BENCH_PAIRS", "96"))
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model = XFM(bench.base_config(), init=bench.gpu_init(dev, 0), device=dev).train()
opt = FlatAdamW(model, lr=1e-4, weight_decay=0.01, lr_mult=2.0)
acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0))
b = {k: v.to(dev) for k, v in bench.make_host_batch(B, 40, 15, model.cfg["vocab_size"], 224, 100).items()}


def step():
    out = model(b["image"], b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
                masked_ids=b["masked_ids"], ret_mim_loss=True, data_source="image")
    loss = out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + out["loss_mim"]
    acc.backward_step(loss, opt)
    acc.optimizer_step(opt, model)
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
for i in range(4):
    n0 = L.launch_count()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"iter {i}: host issue {1e3 * (t1 - t0):.1f} ms, until GPU idle {1e3 * (t2 - t0):.1f} ms, launches {L.launch_count() - n0}", flush=True)
if os.environ.get("XFM_PYPROF"):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
