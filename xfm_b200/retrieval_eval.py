"""Retrieval evaluation: the re-rank loop of Retrieval.py:76-184 (`evaluation`) and the recall metrics
of Retrieval.py:186-231 (`itm_eval`) on xfm_b200.XFMBase.

The reference scores one image (or one text) per python iteration: k_test candidates from the ITC similarity are pushed
through the 12-layer fusion encoder with the image tokens `.repeat(k, 1, 1)`-ed (Retrieval.py:139,156), i.e. the
cross-attention K/V projection of the SAME image is recomputed k = 256 times per row.  Here several rows are scored per
fusion pass and every text sample carries an index to the image whose K/V it attends to (the `kv_index` of the pre-training
ITM pass): image -> text scores one K/V projection per image, text -> image scores one per DISTINCT candidate image of
the group.  Results are the reference's: same top-k candidates, same `-100` fill, same rank sharding
(`step = n // world + 1`) and the same SUM all-reduce of the two score matrices.
"""
import numpy as np
import torch

from . import encoders as E
from . import lib as L
from .xfm import _twin


def _world():
    dist = torch.distributed
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


@torch.no_grad()
def encode_texts(model, text_ids, text_atts, batch=256):
    """Retrieval.py:90-113: text-encoder states (f32 and their bf16 twins) and the normalised ITC features."""
    f32, f16, emb = [], [], []
    for i in range(0, text_ids.shape[0], batch):
        h = model.get_text_embeds(text_ids[i:i + batch], text_atts[i:i + batch])
        f32.append(h)
        f16.append(_twin(h))
        emb.append(model.get_features(text_embeds=h))
    return torch.cat(f32), torch.cat(f16), torch.cat(emb)


@torch.no_grad()
def encode_images(model, image_batches):
    """Retrieval.py:115-127: vision-encoder tokens (bf16 — they are only ever MMA operands downstream) and ITC features."""
    f16, emb = [], []
    for image in image_batches:
        h, _ = model.get_vision_embeds(image)
        f16.append(_twin(h))
        emb.append(model.get_features(image_embeds=h))
    return torch.cat(f16), torch.cat(emb)


def _fusion_scores(model, text32, text16, text_atts, img16, kv_index):
    """itm_head(fusion(text | image)[:, 0])[:, 1] for text samples that attend to image row kv_index[sample]."""
    Bt, Lt, _ = text16.shape
    kmask = E.RobertaStack.additive_mask(text_atts)
    _, out32, _ = model._fusion_run(text16.contiguous(), Bt, Lt, kmask, img16, img16.shape[0], kv_index, False,
                                    text32=text32.contiguous())
    cls = out32.view(Bt, Lt, -1)[:, 0, :].contiguous()
    return model.itm_head(cls)[:, 1]


@torch.no_grad()
def rerank(model, image_feats16, image_embeds, text_feats32, text_feats16, text_embeds, text_atts, k_test,
           pairs_per_pass=2048, shard=None):
    """Retrieval.py:129-176.  Returns (score_matrix_i2t [I, T], score_matrix_t2i [T, I]) as f32 device tensors, already
    summed over ranks when torch.distributed is initialised.  shard=(rank, world) scores that rank's rows only and skips
    the all-reduce (tests, or a caller that reduces elsewhere)."""
    model._prep()
    dev = image_embeds.device
    n_img, n_txt = image_embeds.shape[0], text_embeds.shape[0]
    k = min(k_test, n_txt)
    rank, world = _world() if shard is None else shard
    sims = L.sgemm_f32(image_embeds.float().contiguous(), text_embeds.float().contiguous())   # Retrieval.py:129 (fp32 exact)
    group = max(1, pairs_per_pass // k)

    score_i2t = torch.full((n_img, n_txt), -100.0, device=dev)
    step = n_img // world + 1
    start, end = rank * step, min(n_img, rank * step + step)
    for i0 in range(start, end, group):
        i1 = min(end, i0 + group)
        g = i1 - i0
        topk_idx = sims[i0:i1].topk(k=k, dim=1).indices                     # [g, k] candidate texts per image
        flat = topk_idx.reshape(-1)
        kv_index = torch.arange(g, device=dev, dtype=torch.int32).repeat_interleave(k)
        score = _fusion_scores(model, text_feats32[flat], text_feats16[flat], text_atts[flat], image_feats16[i0:i1], kv_index)
        score_i2t[i0:i1].scatter_(1, topk_idx, score.view(g, k))

    k2 = min(k_test, n_img)
    group = max(1, pairs_per_pass // k2)
    sims_t = sims.t()
    score_t2i = torch.full((n_txt, n_img), -100.0, device=dev)
    step = n_txt // world + 1
    start, end = rank * step, min(n_txt, rank * step + step)
    for j0 in range(start, end, group):
        j1 = min(end, j0 + group)
        g = j1 - j0
        topk_idx = sims_t[j0:j1].topk(k=k2, dim=1).indices                  # [g, k2] candidate images per text
        uniq, inverse = torch.unique(topk_idx.reshape(-1), return_inverse=True)   # K/V once per distinct image
        rows = torch.arange(j0, j1, device=dev).repeat_interleave(k2)
        score = _fusion_scores(model, text_feats32[rows], text_feats16[rows], text_atts[rows], image_feats16[uniq],
                               inverse.to(torch.int32))
        score_t2i[j0:j1].scatter_(1, topk_idx, score.view(g, k2))

    if world > 1 and shard is None:
        torch.distributed.barrier()
        torch.distributed.all_reduce(score_i2t, op=torch.distributed.ReduceOp.SUM)
        torch.distributed.all_reduce(score_t2i, op=torch.distributed.ReduceOp.SUM)
    return score_i2t, score_t2i


@torch.no_grad()
def evaluation(model, image_batches, text_ids, text_atts, config):
    """`evaluation(model, data_loader, tokenizer, device, config)` of Retrieval.py:76 with the tokenizer applied by the
    caller (text_ids / text_atts [T, max_tokens]) and `image_batches` an iterable of image tensors.  Returns the two numpy
    score matrices the reference returns."""
    was_training = model.training
    model.eval()
    t32, t16, temb = encode_texts(model, text_ids, text_atts, config.get("batch_size_test_text", 256))
    i16, iemb = encode_images(model, image_batches)
    s_i2t, s_t2i = rerank(model, i16, iemb, t32, t16, temb, text_atts, config["k_test"])
    model.train(was_training)
    return s_i2t.cpu().numpy(), s_t2i.cpu().numpy()


def itm_eval(scores_i2t, scores_t2i, txt2img, img2txt):
    """Retrieval.py:186-231: recall@{1,5,10} both ways.  Rank of a candidate = its position in the descending argsort of
    the score row (numpy's argsort order for ties, reversed, as in the reference)."""
    def position(score_row, targets):
        inds = np.argsort(score_row)[::-1]
        pos = np.empty_like(inds)
        pos[inds] = np.arange(inds.size)
        return pos[np.asarray(targets)].min()

    ranks = np.array([position(s, img2txt[i]) for i, s in enumerate(scores_i2t)], dtype=np.float64)
    tr1, tr5, tr10 = (100.0 * float((ranks < n).sum()) / len(ranks) for n in (1, 5, 10))
    ranks = np.array([position(s, [txt2img[i]]) for i, s in enumerate(scores_t2i)], dtype=np.float64)
    ir1, ir5, ir10 = (100.0 * float((ranks < n).sum()) / len(ranks) for n in (1, 5, 10))
    tr_mean, ir_mean = (tr1 + tr5 + tr10) / 3, (ir1 + ir5 + ir10) / 3
    return {"txt_r1": tr1, "txt_r5": tr5, "txt_r10": tr10, "txt_r_mean": tr_mean, "img_r1": ir1, "img_r5": ir5,
            "img_r10": ir10, "img_r_mean": ir_mean, "r_mean": (tr_mean + ir_mean) / 2}
