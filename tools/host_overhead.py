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
    pstats.Stats(pr).sort_stats("tottime").print_stats(40)

# ---- does running un-synchronised (host ahead of the GPU) change the step time?
def loop(n, sync_each):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        l = step()
        if sync_each:
            float(l.detach())
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1e3 * (time.perf_counter() - t0) / n


if os.environ.get("XFM_PRERESERVE"):
    gib = int(os.environ["XFM_PRERESERVE"])
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    x = torch.empty(gib << 30, dtype=torch.uint8, device=dev)
    del x
    print("pre-reserved", gib, "GiB; reserved now", torch.cuda.memory_reserved() / 2**30)
for mode in (False, True, False, True, False, True):
    d0 = torch.cuda.memory_stats().get("num_device_alloc")
    ms, wall = loop(6, mode)
    st = torch.cuda.memory_stats()
    print(f"sync_each={mode}: {ms:.1f} ms/step (events), {wall:.1f} ms/step (wall); device_allocs +{st.get('num_device_alloc') - d0}, "
          f"reserved={torch.cuda.memory_reserved() / 2**30:.1f} GiB, peak_alloc={torch.cuda.max_memory_allocated() / 2**30:.1f} GiB",
          flush=True)
