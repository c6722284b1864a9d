"""Throughput of the VQA fine-tuning step and of the answer-ranking inference (BASELINE config #5;
models/model_generation.py:93-202) on one B200: XFM-base + 12-layer causal decoder, random init, synthetic data.

    python tools/bench_vqa.py [--res 384 --batch 24 --answers 3 --answer-len 8 --steps 10] [--cpu-sample]

One JSON line (also appended to gpurun_out/vqa.jsonl):
  train_questions_per_s   questions / s of forward + backward + clip + AdamW (CUDA events, after warm-up)
  rank_questions_per_s    questions / s of forward(train=False): 128-candidate first-token pass + k_test = 128 re-rank
  cpu_questions_per_s     (--cpu-sample) the oracle's training forward + backward on the host cores, bounded sample
"""
import argparse
import json
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=384)
    ap.add_argument("--batch", type=int, default=24)
    ap.add_argument("--tokens", type=int, default=40)
    ap.add_argument("--answers", type=int, default=3)
    ap.add_argument("--answer-len", type=int, default=8)
    ap.add_argument("--candidates", type=int, default=3128)
    ap.add_argument("--k-test", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--cpu-sample", action="store_true")
    a = ap.parse_args()
    import bench as Bn
    from xfm_b200.accelerator import B200DDPAccelerator, FlatAdamW
    from xfm_b200.model_generation import XFMForVQA

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = Bn.base_config()
    cfg.update(image_res=a.res, use_vision_tokenizer=False, num_dec_layers=12, decoder_fusion_start_at=0, pad_token_id=1)
    model = XFMForVQA(cfg, init=Bn.gpu_init(dev, 0), device=dev).train()
    opt = FlatAdamW(model, lr=2e-5, weight_decay=0.01, lr_mult=2.0)
    acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0))
    g = torch.Generator().manual_seed(1)
    V, B, L, La = 50265, a.batch, a.tokens, a.answer_len

    def text(n, length, min_real):
        ids = torch.randint(3, V - 1, (n, length), generator=g)
        ids[:, 0] = 0
        n_real = torch.randint(min_real, length + 1, (n,), generator=g)
        pad = torch.arange(length).view(1, -1) >= n_real.view(-1, 1)
        atts = torch.ones(n, length, dtype=torch.long)
        atts[pad] = 0
        ids[pad] = 1
        return types.SimpleNamespace(input_ids=ids.to(dev), attention_mask=atts.to(dev))

    image = torch.rand(B, 3, a.res, a.res, generator=g).to(dev)
    q = text(B, L, L // 2)
    k = [a.answers] * B
    ans = text(B * a.answers, La, 3)
    weights = (torch.rand(B * a.answers, generator=g) * 0.8 + 0.2).to(dev)
    cand = text(a.candidates, La, 3)

    def step():
        loss = model(image, q, ans, k=k, weights=weights, train=True)
        acc.backward_step(loss, opt)
        acc.optimizer_step(opt, model)
        return loss

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1) * 1e-3 / n

    for _ in range(3):
        step()
    loss, t_train = timed(step, a.steps)
    model.eval()
    with torch.no_grad():
        model(image, q, cand, k=a.k_test, train=False)
        _, t_rank = timed(lambda: model(image, q, cand, k=a.k_test, train=False), 3)
    out = {"what": "VQA fine-tune step / answer ranking (model_generation.py:93-202)", "res": a.res, "questions": B,
           "answers_per_question": a.answers, "answer_len": La, "candidates": a.candidates, "k_test": a.k_test,
           "train_ms_per_step": round(t_train * 1e3, 2), "train_questions_per_s": round(B / t_train, 1),
           "rank_ms": round(t_rank * 1e3, 2), "rank_questions_per_s": round(B / t_rank, 1), "loss": round(float(loss), 4)}
    if a.cpu_sample:   # oracle (test infrastructure) on the host cores, bounded sample: 2 questions, fwd + bwd
        from oracle import xfm_oracle as O
        ocfg = O.base_config(image_res=a.res, dec_layers=12, use_bbox=False)
        sd = O.make_state_dict(ocfg, 0)
        for v in sd.values():
            v.requires_grad_(True)
        b = O.make_vqa_batch(ocfg, B=2, L=L, La=La, n_cand=4, seed=5)
        torch.set_num_threads(os.cpu_count())
        t0 = time.perf_counter()
        O.vqa_train_loss(b["image"], b["q_ids"], b["q_atts"], b["a_ids"], b["a_atts"], b["k"], b["weights"], sd, ocfg).backward()
        dt = time.perf_counter() - t0
        out.update(cpu_questions_per_s=round(2 / dt, 3), cpu_cores=os.cpu_count(),
                   cpu_sample="oracle fwd + bwd, 2 questions / 3 answers, one step")
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/vqa.jsonl", "a") as f:
        f.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
