"""Throughput of the retrieval and NLVR2 fine-tuning steps (BASELINE configs #3 and #4; models/model_retrieval.py:26-37,
models/model_nlvr.py:28-44) on one B200: XFM-base, random init, synthetic data, fwd + bwd + clip + AdamW.

    python tools/bench_finetune.py --task retrieval [--res 384 --batch 32]
    python tools/bench_finetune.py --task nlvr      [--res 384 --batch 64]

One JSON line per run (also appended to gpurun_out/finetune.jsonl): samples / s (CUDA events, after warm-up) and the
algorithmic TFLOP/s from SURVEY.md §8d's per-sample figures (retrieval 579.36, NLVR 836.07 GFLOP fwd + bwd at 384 px).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GFLOP_384 = {"retrieval": 579.36, "nlvr": 836.07}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", required=True, choices=["retrieval", "nlvr"])
    ap.add_argument("--res", type=int, default=384)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--tokens", type=int, default=40)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    import bench as Bn
    from xfm_b200.accelerator import B200DDPAccelerator, FlatAdamW

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B = a.batch or (32 if a.task == "retrieval" else 64)
    cfg = Bn.base_config()
    cfg.update(image_res=a.res, use_vision_tokenizer=False)
    if a.task == "retrieval":
        from xfm_b200.model_retrieval import XFMForRetrieval as Model
    else:
        from xfm_b200.model_nlvr import XFMForNLVR as Model
    model = Model(cfg, init=Bn.gpu_init(dev, 0), device=dev).train()
    opt = FlatAdamW(model, lr=3e-5, weight_decay=0.01, lr_mult=2.0)
    acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0))
    g = torch.Generator().manual_seed(1)
    V, L = 50265, a.tokens
    ids = torch.randint(3, V - 1, (B, L), generator=g)
    ids[:, 0] = 0
    n_real = torch.randint(L // 2, L + 1, (B,), generator=g)
    pad = torch.arange(L).view(1, -1) >= n_real.view(-1, 1)
    atts = torch.ones(B, L, dtype=torch.long)
    atts[pad] = 0
    ids[pad] = 1
    ids, atts = ids.to(dev), atts.to(dev)
    n_img = B if a.task == "retrieval" else 2 * B
    image = torch.rand(n_img, 3, a.res, a.res, generator=g).to(dev)
    idx = torch.randint(0, max(1, B // 5 * 4), (B,), generator=g).to(dev)     # 5 captions per image => duplicates
    targets = torch.randint(0, 2, (B,), generator=g).to(dev)

    def step():
        if a.task == "retrieval":
            l_itc, l_itm = model(image, ids, atts, idx=idx)
            loss = l_itc + l_itm
        else:
            loss = model(image, ids, atts, targets)
        acc.backward_step(loss, opt)
        acc.optimizer_step(opt, model)
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / a.steps
    out = {"what": f"{a.task} fine-tune step", "res": a.res, "samples": B, "images": n_img, "ms_per_step": round(t * 1e3, 2),
           "samples_per_s": round(B / t, 1), "loss": round(float(loss.detach()), 4)}
    if a.res == 384:
        out["algorithmic_tflops"] = round(GFLOP_384[a.task] * B / t / 1e3, 1)
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/finetune.jsonl", "a") as f:
        f.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
