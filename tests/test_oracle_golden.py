"""Pin the oracle (oracle/xfm_oracle.py) against fixtures produced by the UNMODIFIED reference
(tools/make_golden.py via oracle/ref_shim.py).  CPU only."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import xfm_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _masks(cfg, B, seed):
    random.seed(seed)
    np.random.seed(seed)
    return O.sample_mim_masks(cfg, B)


def test_masking_generator_bit_exact(golden_dir):
    gold = _load(golden_dir, "masks.pt")
    assert len(gold) == 9
    for (size, n, mn, seed), ref in gold.items():
        random.seed(seed)
        np.random.seed(seed)
        gen = O.MaskingGenerator(size, num_masking_patches=n, min_num_patches=mn)
        mine = torch.from_numpy(np.stack([gen() for _ in range(6)]))
        assert torch.equal(mine, ref), (size, n, mn, seed)
        assert int(mine[0].sum()) == n


def test_itc_with_idx_and_hard_negative_weights(golden_dir):
    g = _load(golden_dir, "itc_idx.pt")
    temp = torch.tensor(g["temp"])
    li = O.contrastive_loss(g["image_feat"], g["text_feat"], temp, idx_all=g["idx"])
    lp = O.contrastive_loss(g["image_feat"], g["text_feat"], temp)
    assert abs(float(li) - g["loss_idx"]) < 1e-6
    assert abs(float(lp) - g["loss_plain"]) < 1e-6
    w_i2t, w_t2i = O.hard_negative_weights(g["image_feat"], g["text_feat"], temp, idx=g["idx"])
    torch.testing.assert_close(w_i2t, g["weights_i2t"], rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(w_t2i, g["weights_t2i"], rtol=1e-6, atol=1e-8)


def _run_oracle(g, want_grads=False):
    cfg = g["cfg"]
    sd = O.make_state_dict(cfg, seed=0)
    if want_grads:
        for v in sd.values():
            v.requires_grad_(True)
    batch = O.make_batch(cfg, g["B"], L=g["L"], M=g["M"], seed=1, image_uniform=g["image_uniform"])
    ids_mask = _masks(cfg, g["B"], g["mask_seed"])
    col = {}
    with torch.set_grad_enabled(want_grads):
        out = O.pretrain_forward(sd, cfg, batch, g["image_neg_idx"], g["text_neg_idx"], ids_mask=ids_mask, collect=col)
    return sd, out, col, ids_mask


@pytest.mark.parametrize("name", ["tiny_vq.pt", "tiny_mse.pt"])
def test_tiny_losses_masks_and_grads(golden_dir, name):
    g = _load(golden_dir, name)
    sd, out, col, ids_mask = _run_oracle(g, want_grads=True)
    assert torch.equal(ids_mask, g["ids_mask"])  # MIM masks: bit-exact
    for k, v in g["losses"].items():
        assert abs(float(out[k]) - v) <= 2e-5 * max(1.0, abs(v)), (k, float(out[k]), v)
    torch.testing.assert_close(col["weights_i2t"], g["weights_i2t"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(col["weights_t2i"], g["weights_t2i"], rtol=1e-5, atol=1e-7)
    assert torch.equal(torch.argmax(col["weights_t2i"], 1), g["image_neg_idx"])
    total = out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + out["loss_mim"]
    total.backward()
    tied = {"fusion_encoder.roberta.embeddings.word_embeddings.weight", "text_encoder.roberta.embeddings.word_embeddings.weight"}
    for n, gr in g["grads"].items():
        mine = sd[n].grad
        assert mine is not None, n
        torch.testing.assert_close(mine, gr, rtol=2e-4, atol=2e-6, msg=lambda m, n=n: f"{n}: {m}")
    # parameters the reference leaves without a gradient must not receive one from the oracle either
    for n in g["grad_none"]:
        if n in sd:
            assert sd[n].grad is None or float(sd[n].grad.abs().max()) == 0.0, n


def test_tiny_vq_activations_and_ids(golden_dir):
    g = _load(golden_dir, "tiny_vq.pt")
    sd, out, col, _ = _run_oracle(g)
    for grp in ("vision", "text", "fusion_pos", "vision_masked"):
        assert len(col[grp]) == len(g["acts"][grp])
        for i, (a, b) in enumerate(zip(col[grp], g["acts"][grp])):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=2e-5, msg=lambda m, i=i, grp=grp: f"{grp}[{i}]: {m}")
    for k in ("image_embeds", "text_embeds", "image_embeds_masked", "image_feat", "text_feat"):
        torch.testing.assert_close(col[k], g[k], rtol=1e-4, atol=2e-5)
    cfg = g["cfg"]
    batch = O.make_batch(cfg, g["B"], L=g["L"], M=g["M"], seed=1, image_uniform=True)
    ids = O.vqkd_codebook_indices(batch["image"], sd, cfg)
    assert ids.dtype == torch.int64 and ids.shape == g["vq_ids"].shape
    assert torch.equal(ids, g["vq_ids"])  # VQ-KD token ids: bit-exact
    assert ids.unique().numel() > 8  # not collapsed


@pytest.mark.parametrize("name", ["base_mse.pt", "base_vq.pt"])
def test_base_config_against_reference(golden_dir, name):
    path = os.path.join(golden_dir, name)
    if not os.path.exists(path):
        pytest.skip("base fixture not generated")
    g = _load(golden_dir, name)
    torch.set_num_threads(os.cpu_count())
    sd, out, col, ids_mask = _run_oracle(g)
    assert torch.equal(ids_mask, g["ids_mask"])
    for k, v in g["losses"].items():
        assert abs(float(out[k]) - v) <= 5e-5 * max(1.0, abs(v)), (k, float(out[k]), v)
    tok, nd = [0, 1, 7, -1], 16
    for grp in ("vision", "text", "fusion_pos", "vision_masked"):
        for i, (a, s) in enumerate(zip(col[grp], g["acts_summary"][grp])):
            torch.testing.assert_close(a[:, tok, :nd], s["slice"], rtol=2e-4, atol=1e-4)
            assert abs(float(a.abs().max()) - s["absmax"]) <= 1e-3 * max(1.0, s["absmax"])
    if "vq_ids" in g:
        cfg = g["cfg"]
        batch = O.make_batch(cfg, g["B"], L=g["L"], M=g["M"], seed=1, image_uniform=True)
        ids = O.vqkd_codebook_indices(batch["image"], sd, cfg)
        amb = O.quantizer_ambiguous(O.vqkd_features(O.vqkd_preprocess(batch["image"]), sd, cfg),
                                    sd["vqkd.quantize.embedding.weight"]).view(ids.shape)
        assert torch.equal(ids[~amb], g["vq_ids"][~amb])


def test_vqa_decoder_against_reference(golden_dir):
    """XFMForVQA (models/model_generation.py:93-202): training loss, per-answer losses, question states, gradients and the
    rank_answer re-ranking of the oracle against the unmodified reference."""
    g = _load(golden_dir, "tiny_vqa.pt")
    cfg = g["cfg"]
    sd = O.make_state_dict(cfg, seed=0)
    for v in sd.values():
        v.requires_grad_(True)
    b = O.make_vqa_batch(cfg)
    col = {}
    loss = O.vqa_train_loss(b["image"], b["q_ids"], b["q_atts"], b["a_ids"], b["a_atts"], b["k"], b["weights"], sd, cfg,
                            collect=col)
    assert abs(float(loss) - g["loss"]) <= 2e-5 * abs(g["loss"])
    torch.testing.assert_close(col["answer_loss"], g["answer_loss"], rtol=2e-5, atol=1e-5)
    torch.testing.assert_close(col["question_output"], g["question_output"], rtol=1e-4, atol=2e-5)
    loss.backward()
    for n, gr in g["grads"].items():
        torch.testing.assert_close(sd[n].grad, gr, rtol=2e-4, atol=2e-6, msg=lambda m, n=n: f"{n}: {m}")
    with torch.no_grad():
        ids, probs = O.vqa_rank(b["image"], b["q_ids"], b["q_atts"], b["cand_ids"], b["cand_atts"], g["k_test"], sd, cfg)
    assert torch.equal(ids, g["topk_ids"])
    torch.testing.assert_close(probs, g["topk_probs"], rtol=1e-4, atol=1e-7)


def test_region_branch_against_reference(golden_dir):
    """Region / bbox branch (model_pretrain.py:39-41,81-86; xfm.py:574-597,815-854; beit2.py:468-475; box_ops.py): the five
    losses, region-pooled embeddings, predicted boxes and gradients of the restatement equal the reference's."""
    g = _load(golden_dir, "tiny_region.pt")
    cfg = g["cfg"]
    sd = O.make_state_dict(cfg, seed=0)
    for v in sd.values():
        if v.dtype.is_floating_point:
            v.requires_grad_(True)
    batch = O.make_region_batch(cfg)
    col = {}
    out = O.pretrain_forward_region(sd, cfg, batch, g["image_neg_idx"], g["text_neg_idx"], collect=col)
    for k, v in out.items():
        assert abs(float(v) - g["losses"][k]) <= 2e-5 * max(1.0, abs(g["losses"][k])), (k, float(v), g["losses"][k])
    assert g["losses"]["loss_mim"] == 0.0          # no MIM on region batches (model_pretrain.py:67)
    torch.testing.assert_close(col["image_embeds"], g["image_embeds"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(col["image_embeds_fullatts"], g["image_embeds_fullatts"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(col["output_coord"], g["output_coord"], rtol=1e-5, atol=1e-6)
    sum(out.values()).backward()
    for n, ref in g["grads"].items():
        assert float((sd[n].grad - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-9, n
    # degenerate boxes zero the GIoU term for the whole batch (xfm.py:825-828); is_image rows are excluded from both terms
    coord = torch.tensor([[0.5, 0.5, 0.2, 0.2], [0.3, 0.3, -0.1, 0.2]])
    lb, lg = O.bbox_loss(coord, torch.tensor([[0.5, 0.5, 0.2, 0.2], [0.3, 0.3, 0.1, 0.2]]))
    assert float(lg) == 0.0 and abs(float(lb) - 0.1) < 1e-7


def test_bert_text_encoder_variant_against_reference(golden_dir):
    """models/xbert.py as the text encoder: BertForMaskedLM naming, absolute position ids, and the
    placement of 1/sqrt(d) (xbert.py:296-301,329-330) — the reference produced bit-identical results for both placements."""
    g = _load(golden_dir, "tiny_bert.pt")
    assert g["scale_order_identical"]
    cfg = g["cfg"]
    assert cfg["text_arch"] == "bert"
    batch = O.make_batch(cfg, 4, L=24, M=6, seed=1)
    for fp16 in (True, False):
        sd = O.make_state_dict(cfg, seed=0)
        for v in sd.values():
            if v.dtype.is_floating_point:
                v.requires_grad_(True)
        c = dict(cfg, text_fp16=fp16)
        out = O.pretrain_forward(sd, c, batch, g["image_neg_idx"], g["text_neg_idx"])
        for k, v in g["losses"].items():
            assert abs(float(out[k]) - v) <= 2e-5 * max(1.0, abs(v)), (fp16, k, float(out[k]), v)
        mlm = O.text_mlm_loss(batch["text_ids_masked"], batch["text_atts"], batch["masked_pos"], batch["masked_ids"], sd, c)
        assert abs(float(mlm) - g["text_only_mlm"]) <= 2e-5 * abs(g["text_only_mlm"])
        torch.testing.assert_close(O.text_forward(batch["text_ids"], batch["text_atts"], sd, c).detach(), g["text_embeds"],
                                   rtol=1e-5, atol=1e-5)
        (out["loss_itc"] + out["loss_itm"] + out["loss_mlm"]).backward()
        for n, ref in g["grads"].items():
            assert float((sd[n].grad - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-9, (fp16, n)


def test_finetune_forwards_against_reference(golden_dir):
    """BASELINE configs #3 / #4 at tiny widths: the oracle pieces, composed the way tests/test_model_gpu.py composes them for
    the retrieval and NLVR models, equal the reference's OWN models/model_retrieval.py:26-37 and models/model_nlvr.py:28-44
    forwards (tiny_finetune.pt: losses, prediction, gradients incl. the text path opened by is_pretrain=False)."""
    g = _load(golden_dir, "tiny_finetune.pt")
    cfg = O.tiny_config()

    def fresh_sd():
        sd = O.make_state_dict(cfg, seed=0)
        for v in sd.values():
            if v.dtype.is_floating_point:
                v.requires_grad_(True)
        return sd

    def close(a, b, what):
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-9, what

    # retrieval: ITC with idx soft labels + idx-masked hard negatives (the heaviest one, as the fixture's multinomial stand-in)
    r, sd = g["retrieval"], fresh_sd()
    batch = O.make_batch(cfg, 6, L=24, M=6, seed=3)
    ie = O.vision_forward(batch["image"], sd, cfg)
    ia = torch.ones(ie.shape[:2], dtype=torch.long)
    te = O.text_forward(batch["text_ids"], batch["text_atts"], sd, cfg)
    fi, ft = O.get_features(ie, te, sd)
    w_i2t, w_t2i = O.hard_negative_weights(fi.detach(), ft.detach(), sd["temp"].detach(), r["idx"])
    tneg, ineg = w_i2t.argmax(1), w_t2i.argmax(1)
    itc = O.contrastive_loss(fi, ft, sd["temp"], r["idx"])
    itm, _ = O.matching_loss(ie, ia, te, batch["text_atts"], ineg, tneg, sd, cfg, is_pretrain=False)
    assert abs(float(itc) - r["loss_itc"]) <= 2e-5 * abs(r["loss_itc"]), (float(itc), r["loss_itc"])
    assert abs(float(itm) - r["loss_itm"]) <= 2e-5 * abs(r["loss_itm"]), (float(itm), r["loss_itm"])
    (itc + itm).backward()
    for n, ref in r["grads"].items():
        close(sd[n].grad.reshape(ref.shape), ref, n)
    assert float(r["grads"]["text_encoder.roberta.encoder.layer.0.intermediate.dense.weight"].abs().max()) > 0

    # NLVR: two images per text, shared text states, concatenated CLS -> Linear, LayerNorm, GELU, Linear -> CE
    n, sd = g["nlvr"], fresh_sd()
    B = n["targets"].numel()
    head = {k: v.clone().requires_grad_(True) for k, v in n["head"].items()}
    for k, v in head.items():
        assert torch.equal(v.detach(), O.make_tensor("cls_head." + k, tuple(v.shape), 0))
    batch = O.make_batch(cfg, 2 * B, L=24, M=6, seed=5)
    ids, atts = batch["text_ids"][:B], batch["text_atts"][:B]
    ie = O.vision_forward(batch["image"], sd, cfg)
    ia = torch.ones(ie.shape[:2], dtype=torch.long)
    te = O.text_forward(ids, atts, sd, cfg)
    x = torch.cat([O.fusion_forward(te, atts, ie[:B], ia[:B], sd, cfg)[:, 0], O.fusion_forward(te, atts, ie[B:], ia[B:], sd, cfg)[:, 0]], -1)
    x = torch.nn.functional.linear(x, head["0.weight"], head["0.bias"])
    x = torch.nn.functional.gelu(torch.nn.functional.layer_norm(x, (x.shape[-1],), head["1.weight"], head["1.bias"], 1e-5))
    pred = torch.nn.functional.linear(x, head["3.weight"], head["3.bias"])
    loss = torch.nn.functional.cross_entropy(pred, n["targets"])
    assert abs(float(loss) - n["loss"]) <= 2e-5 * abs(n["loss"]), (float(loss), n["loss"])
    close(pred.detach(), n["prediction"], "prediction")
    loss.backward()
    for k, ref in n["grads"].items():
        got = head[k[len("cls_head."):]].grad if k.startswith("cls_head.") else sd[k].grad
        close(got.reshape(ref.shape), ref, k)
