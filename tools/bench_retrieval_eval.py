"""Throughput of the retrieval re-rank loop (Retrieval.py:76-184) on one B200, XFM-base weights
(random init), synthetic data: N_IMG images, N_TXT texts, k_test candidates per row.

    python tools/bench_retrieval_eval.py [--images 256 --texts 1280 --k 128 --res 224] [--cpu-sample]

Reports (one JSON line, also appended to gpurun_out/retrieval_eval.jsonl):
  pairs_per_s            (image, text) pairs scored per second by xfm_b200.retrieval_eval.rerank (both directions), device time
  per_row_loop_pairs_per_s   the reference's schedule on the same kernels: one row per fusion pass with the image tokens
                             repeated k times (no K/V sharing), i.e. what switching modules alone would give
  encode_s               text + image encoder time
  cpu_pairs_per_s        (--cpu-sample) the oracle's per-row loop on the host cores for a bounded sample
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--texts", type=int, default=1280)
    ap.add_argument("--k", type=int, default=128)
    ap.add_argument("--res", type=int, default=224)
    ap.add_argument("--tokens", type=int, default=40)
    ap.add_argument("--cpu-sample", action="store_true")
    a = ap.parse_args()
    import bench as Bn
    from xfm_b200 import retrieval_eval as RE
    from xfm_b200.model_retrieval import XFMForRetrieval

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = Bn.base_config()
    cfg.update(image_res=a.res, use_vision_tokenizer=False)
    model = XFMForRetrieval(cfg, init=Bn.gpu_init(dev, 0), device=dev).eval()
    g = torch.Generator().manual_seed(1)
    vocab = 50265
    ids = torch.randint(3, vocab - 1, (a.texts, a.tokens), generator=g)
    ids[:, 0] = 0
    n_real = torch.randint(a.tokens // 2, a.tokens + 1, (a.texts,), generator=g)
    pad = torch.arange(a.tokens).view(1, -1) >= n_real.view(-1, 1)
    atts = torch.ones(a.texts, a.tokens, dtype=torch.long)
    atts[pad] = 0
    ids[pad] = 1
    ids, atts = ids.to(dev), atts.to(dev)
    images = [torch.rand(min(64, a.images - i), 3, a.res, a.res, generator=g).to(dev) for i in range(0, a.images, 64)]

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1) * 1e-3

    with torch.no_grad():
        RE.encode_texts(model, ids[:256], atts[:256])   # warm-up (lazy init, allocator)
        (t32, t16, temb), t_txt = timed(lambda: RE.encode_texts(model, ids, atts))
        (i16, iemb), t_img = timed(lambda: RE.encode_images(model, images))
        RE.rerank(model, i16[:8], iemb[:8], t32, t16, temb, atts, a.k)   # warm-up
        _, t_rr = timed(lambda: RE.rerank(model, i16, iemb, t32, t16, temb, atts, a.k))
        pairs = a.images * min(a.k, a.texts) + a.texts * min(a.k, a.images)

        # the reference's schedule on the same kernels: one row per pass, image tokens repeated k times
        rows = min(16, a.images)
        sims = iemb @ temb.t()

        def per_row():
            for i in range(rows):
                idx = sims[i].topk(a.k).indices
                enc = i16[i].repeat(a.k, 1, 1)
                RE._fusion_scores(model, t32[idx], t16[idx], atts[idx], enc, None)
        per_row()
        _, t_row = timed(per_row)
    out = {"what": "retrieval re-rank (Retrieval.py:76-184)", "images": a.images, "texts": a.texts, "k_test": a.k, "res": a.res,
           "pairs": pairs, "rerank_s": round(t_rr, 4), "pairs_per_s": round(pairs / t_rr, 1),
           "per_row_loop_pairs_per_s": round(rows * a.k / t_row, 1), "encode_s": round(t_txt + t_img, 4)}

    if a.cpu_sample:   # oracle per-row loop (test infrastructure) on the host cores, bounded sample
        from oracle import xfm_oracle as O
        ocfg = O.base_config(image_res=a.res)
        sd = O.make_state_dict(ocfg, 0)
        kk = min(a.k, 16)
        bt = O.make_batch(ocfg, kk, L=a.tokens, M=1, seed=2)
        bi = O.make_batch(ocfg, 1, L=a.tokens, M=1, seed=3)
        torch.set_num_threads(os.cpu_count())
        with torch.no_grad():
            ie = O.vision_forward(bi["image"], sd, ocfg)
            te = O.text_forward(bt["text_ids"], bt["text_atts"], sd, ocfg)
            t0 = time.perf_counter()
            enc = ie[0].repeat(kk, 1, 1)
            o = O.fusion_forward(te, bt["text_atts"], enc, torch.ones(enc.shape[:2], dtype=torch.long), sd, ocfg)
            O.itm_head(o[:, 0, :], sd)
            dt = time.perf_counter() - t0
        out.update(cpu_pairs_per_s=round(kk / dt, 2), cpu_cores=os.cpu_count(), cpu_sample=f"1 image x {kk} texts, oracle fusion + itm_head")
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/retrieval_eval.jsonl", "a") as f:
        f.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
