"""Worker of tests/test_dist_gpu.py (torchrun --nproc-per-node W tests/ddp_worker.py): the overlapped gradient all-reduce
(non-vision ranges at the start of the last backward node, vision blocks as they finish, on a side stream) must give the
same parameters as the plain post-backward all-reduce, every rank the same ones, also with two backward passes per
optimizer step (gradient accumulation, Pretrain.py:218-243) in "auto" and in forced-overlap (stash) mode."""
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from xfm_b200.accelerator import B200DDPAccelerator, FlatAdamW  # noqa: E402
from xfm_b200.model_pretrain import XFM  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 8


def run(overlap, accumulate=1, steps=3):
    random.seed(5 + rank)
    np.random.seed(5 + rank)
    torch.manual_seed(5 + rank)
    model = XFM(bench.base_config(), init=bench.gpu_init(dev, 0), device=dev).train()
    model._seed = 77 + rank
    opt = FlatAdamW(model, lr=1e-3, weight_decay=0.01, lr_mult=2.0)
    acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0, OVERLAP_ALLREDUCE=overlap))
    wrapped, opt, _ = acc.set_up(model, opt, None, local, world, rank)
    bs = [{k: v.to(dev) for k, v in bench.make_host_batch(B, 40, 15, model.cfg["vocab_size"], 224, 100 + rank + 7 * j).items()}
          for j in range(accumulate)]
    early = 0
    for _ in range(steps):
        for b in bs:
            out = wrapped(b["image"], b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"],
                          masked_pos=b["masked_pos"], masked_ids=b["masked_ids"], ret_mim_loss=True, data_source="image")
            loss = out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + out["loss_mim"]
            acc.backward_step(loss, opt)
        early += int(len(acc._early) > 0)
        acc.optimizer_step(opt, wrapped)
    torch.cuda.synchronize()
    return model.flat.P[:acc._train_end].clone(), early, float(loss)


def check(tag, overlap, accumulate, want_early):
    p1, e1, l1 = run(overlap, accumulate)
    p0, e0, l0 = run(False, accumulate)
    diff = float((p1 - p0).abs().max())
    other = p1.clone()
    dist.broadcast(other, src=0)
    cross = float((p1 - other).abs().max())
    print(f"DDP_CHECK rank {rank} {tag}: early-reduce steps {e1}/3 (off: {e0}); max |P_overlap - P_plain| = {diff:.3e}; "
          f"max |P - P_rank0| = {cross:.3e}; loss {l1:.5f} vs {l0:.5f}", flush=True)
    # replicas must stay bit-identical; overlap vs plain differ only through non-deterministic gradient summation order
    # (fp32 atomics in split-K wgrad / LayerNorm dgamma), which Adam's first steps turn into +-lr flips of near-zero gradients
    assert e1 == want_early and e0 == 0 and cross == 0.0 and abs(l1 - l0) < 5e-3 * abs(l0), (tag, e1, e0, diff, cross, l1, l0)


check("overlap=True, 1 backward/step", True, 1, 3)
check("overlap=auto, 1 backward/step", "auto", 1, 2)    # the first step only learns the pattern
check("overlap=auto, 2 backward/step", "auto", 2, 2)
check("overlap=True, 2 backward/step (stash)", True, 2, 3)
dist.barrier()
dist.destroy_process_group()
