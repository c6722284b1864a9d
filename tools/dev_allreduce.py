"""NCCL all-reduce timing at the gradient-buffer sizes of the step (torchrun, one rank per GPU).
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/dev_allreduce.py
"""
import json
import os

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
N_ALL = 362_000_000   # fp32 elements of the flat gradient buffer (XFM-base + VQ-KD tokenizer excluded)
cases = [("block_28MB_f32", 7_100_000, torch.float32), ("vision_344MB_f32", 86_000_000, torch.float32),
         ("all_1448MB_f32", N_ALL, torch.float32), ("all_724MB_bf16", N_ALL, torch.bfloat16),
         ("half_724MB_f32", N_ALL // 2, torch.float32), ("quarter_362MB_f32", N_ALL // 4, torch.float32)]
for name, n, dt in cases:
    x = torch.ones(n, dtype=dt, device=dev)
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        b = n * x.element_size()
        print(json.dumps(dict(case=name, world=world, bytes=b, ms=round(float(ms), 3), algbw_gbs=round(b / float(ms) / 1e6, 1),
                              busbw_gbs=round(b / float(ms) / 1e6 * 2 * (world - 1) / world, 1))), flush=True)
    del x
dist.barrier()
dist.destroy_process_group()
