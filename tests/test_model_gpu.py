"""Parity of the CUDA path (xfm_b200.XFM through the reference module API) against
  (a) fixtures produced by the UNMODIFIED reference (tests/golden/*.pt, tools/make_golden.py), and
  (b) the CPU oracle (oracle/xfm_oracle.py) on the same seeded inputs.
Tolerances follow BASELINE.json north_star: MIM masks / VQ ids bit-exact, per-layer activations <= 2e-2 max-abs
(bf16 compute vs the fp32 reference), losses <= 1e-3 relative (5e-3 where noted for the tiny VQ-CE head)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import xfm_oracle as O

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _build(g, train=False):
    from xfm_b200.model_pretrain import XFM
    cfg = dict(g["cfg"])
    model = XFM(cfg, init=lambda n, s: O.make_tensor(n, s, 0), device="cuda")
    model.train(train)
    return model, cfg


def _batch(g, cfg):
    b = O.make_batch(cfg, g["B"], L=g["L"], M=g["M"], seed=1, image_uniform=g["image_uniform"])
    return {k: v.cuda() for k, v in b.items()}


def _run(model, g, batch, collect=False):
    model._forced_negatives = (g["image_neg_idx"], g["text_neg_idx"])
    model._forced_masks = g["ids_mask"]
    return model(batch["image"], batch["text_ids"], batch["text_atts"], text_ids_masked=batch["text_ids_masked"],
                 masked_pos=batch["masked_pos"], masked_ids=batch["masked_ids"], ret_mim_loss=True, data_source="image")


def _maxabs(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max())


def test_state_dict_layout_matches_reference_names():
    from xfm_b200.model_pretrain import XFM
    cfg = O.tiny_config(use_vision_tokenizer=True)
    model = XFM(cfg, init=lambda n, s: O.make_tensor(n, s, 0), device="cuda")
    want = O.expand_tied(O.make_state_dict(cfg), cfg)
    have = model.state_dict()
    missing = [k for k in want if k not in have]
    assert not missing, missing[:10]
    for k, v in want.items():
        assert tuple(have[k].shape) == tuple(v.shape), k
        if v.dtype.is_floating_point:
            torch.testing.assert_close(have[k].cpu(), v, rtol=0, atol=0, msg=k)
    # tied weights are the same storage (xroberta.py:1209-1210)
    p = dict(model.named_parameters(remove_duplicate=False))
    assert p["fusion_encoder.lm_head.decoder.weight"] is p["fusion_encoder.roberta.embeddings.word_embeddings.weight"]
    assert set(model.init_params) >= {"temp", "vision_proj.weight", "text_proj.bias", "itm_head.0.weight"}
    # round trip
    sd = {k: v.clone() for k, v in have.items()}
    model.load_state_dict(sd)


def test_tiny_vq_activations_losses_against_reference(golden_dir):
    g = _load(golden_dir, "tiny_vq.pt")
    model, cfg = _build(g)
    batch = _batch(g, cfg)
    model._vis.collect, model._txt.collect, model._fus.collect = [], [], []
    with torch.no_grad():
        out = _run(model, g, batch)
    nv, nt, nf = cfg["vision_depth"], cfg["text_layers"], cfg["fusion_layers"]
    vis, txt, fus = model._vis.collect, model._txt.collect, model._fus.collect
    # call order: vision (ONE 2B-sample pass: clean images, then their masked copies) | text (ONE 2B-sample pass: clean rows,
    # then the masked copies) | fusion (4B pass: 3B ITM rows + B MLM rows)
    assert len(vis) == nv and len(txt) == nt and len(fus) == nf
    B = g["B"]
    for i in range(nv):
        assert vis[i].shape[0] == 2 * B
        assert _maxabs(vis[i][:B], g["acts"]["vision"][i]) <= 2e-2, ("vision", i)
        assert _maxabs(vis[i][B:], g["acts"]["vision_masked"][i]) <= 2e-2, ("vision_masked", i)
    for i in range(nt):
        assert txt[i].shape[0] == 2 * B
        assert _maxabs(txt[i][:B], g["acts"]["text"][i]) <= 2e-2, ("text", i)
        assert _maxabs(txt[i][B:], g["acts"]["text_masked"][i]) <= 2e-2, ("text_masked", i)
    for i in range(nf):  # first B samples of the fusion pass are the positive pairs
        assert _maxabs(fus[i][:B], g["acts"]["fusion_pos"][i]) <= 2e-2, ("fusion", i)
    for k, v in g["losses"].items():
        tol = 5e-3 if k == "loss_mim" else 1e-3
        assert abs(float(out[k]) - v) <= tol * max(1.0, abs(v)), (k, float(out[k]), v)


def test_tiny_features_weights_and_vq_ids(golden_dir):
    g = _load(golden_dir, "tiny_vq.pt")
    model, cfg = _build(g)
    batch = _batch(g, cfg)
    with torch.no_grad():
        ie, ia = model.get_vision_embeds(batch["image"])
        te = model.get_text_embeds(batch["text_ids"], batch["text_atts"])
        fi, ft = model.get_features(ie, te)
        w_i2t, w_t2i = model.hard_negative_weights(fi, ft)
        iem, _, ids_mask = model.get_vision_embeds(batch["image"], do_mask=True) if False else (None, None, None)
        ids = model.get_codebook_indices(batch["image"])
    assert ia.dtype == torch.long and ia.shape == ie.shape[:2] and bool((ia == 1).all())
    assert _maxabs(ie, g["image_embeds"]) <= 2e-2 and _maxabs(te, g["text_embeds"]) <= 2e-2
    assert _maxabs(fi, g["image_feat"]) <= 5e-3 and _maxabs(ft, g["text_feat"]) <= 5e-3
    assert _maxabs(w_i2t, g["weights_i2t"]) <= 2e-2 and _maxabs(w_t2i, g["weights_t2i"]) <= 2e-2
    assert ids.dtype == torch.int64 and ids.shape == g["vq_ids"].shape


def _vq_margins(cfg, B, L, M):
    """fp64 gap between the best and second-best codebook distance of every patch, from the oracle's fp32 z."""
    sd = O.make_state_dict(cfg, 0)
    img = O.make_batch(cfg, B, L=L, M=M, seed=1, image_uniform=True)["image"]
    with torch.no_grad():
        z = O.vqkd_features(O.vqkd_preprocess(img), sd, cfg)
    zf = torch.nn.functional.normalize(z.permute(0, 2, 3, 1), dim=-1).reshape(-1, cfg["codebook_dim"]).double()
    cb = sd["vqkd.quantize.embedding.weight"].double()
    gaps = []
    for i in range(0, zf.shape[0], 4096):
        top2 = torch.topk(O.quantizer_distances(zf[i:i + 4096], cb), 2, dim=1, largest=False).values
        gaps.append(top2[:, 1] - top2[:, 0])
    return torch.cat(gaps)


@pytest.mark.parametrize("name", ["tiny_vq.pt", "base_vq.pt"])
def test_vq_ids_end_to_end_against_reference(golden_dir, record, name):
    """VQKD.get_codebook_indices (model_vqkd.py:173-175) end to end against the ids the UNMODIFIED reference produced.
    The quantizer is bit-exact on identical z (test_kernels_gpu.py::test_vq_argmin_bit_exact) and encode_task_layer runs at
    fp32-grade precision like the reference (model_vqkd.py:154-155); what is left is the bf16 tokenizer ViT (the reference's
    own GPU path runs it under fp16 autocast, xfm.py:627): its perturbation of z (~1e-2) can flip an argmin whose fp64 margin
    is below that.  Every id whose margin exceeds the perturbation must match; the match rate is recorded."""
    g = _load(golden_dir, name)
    model, cfg = _build(g)
    batch = _batch(g, cfg)
    with torch.no_grad():
        ids = model.get_codebook_indices(batch["image"]).cpu()
    want = g["vq_ids"]
    assert ids.shape == want.shape
    gap = _vq_margins(cfg, g["B"], g["L"], g["M"]).view(want.shape)
    match = ids == want
    rates = {f"margin>{t}": float(match[gap > t].float().mean()) for t in (0.0, 1e-3, 1e-2, 3e-2)}
    worst_missed = float(gap[~match].max()) if (~match).any() else 0.0
    record("vq_ids", fixture=name, n=int(want.numel()), match=float(match.float().mean()), worst_missed_margin=worst_missed,
           **rates)
    # measured (profiles/r02_parity_measurements.jsonl): tiny 64 / 64; base (B=8) 98.6 % with the largest margin among the
    # differing ids 3.6e-3; every id whose margin exceeds 1e-2 matches
    assert worst_missed <= 1e-2, worst_missed          # no id with a margin above the bf16 perturbation differs
    assert float(match.float().mean()) >= 0.975


def test_mim_masks_bit_exact_through_the_module(golden_dir):
    g = _load(golden_dir, "tiny_vq.pt")
    model, cfg = _build(g)
    batch = _batch(g, cfg)
    random.seed(g["mask_seed"])
    np.random.seed(g["mask_seed"])
    with torch.no_grad():
        _, _, ids_mask = model.get_vision_embeds(batch["image"], do_mask=True)
    assert ids_mask.dtype == torch.bool
    assert torch.equal(ids_mask.cpu(), g["ids_mask"])


@pytest.mark.parametrize("name", ["tiny_vq.pt", "tiny_mse.pt"])
def test_tiny_losses_and_gradients(golden_dir, name):
    g = _load(golden_dir, name)
    model, cfg = _build(g)
    batch = _batch(g, cfg)
    out = _run(model, g, batch)
    for k, v in g["losses"].items():
        # tiny configs: the VQ-CE head sees bf16-perturbed token ids, and temp=0.07 multiplies feature error by 14 in ITC
        tol = 5e-3 if (k == "loss_mim" and cfg["use_vision_tokenizer"]) else (2e-3 if k == "loss_itc" else 1e-3)
        assert abs(float(out[k]) - v) <= tol * max(1.0, abs(v)), (k, float(out[k]), v)
    total = out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + out["loss_mim"]
    total.backward()
    params = dict(model.named_parameters())
    bad = []
    for n, ref in g["grads"].items():
        mine = params[n].grad
        assert mine is not None, n
        err = _maxabs(mine, ref) / max(float(ref.abs().max()), 1e-8)
        if err > 5e-2:
            bad.append((n, err))
    assert not bad, bad
    # parameters the reference leaves without a gradient have none here either
    for n in g["grad_none"]:
        if n in params and params[n].requires_grad:
            assert params[n].grad is None, n
    # a second backward accumulates; zero_grad(set_to_none) resets
    g1 = params["itm_head.0.weight"].grad.clone()
    out2 = _run(model, g, batch)
    (out2["loss_itc"] + out2["loss_itm"] + out2["loss_mlm"] + out2["loss_mim"]).backward()
    assert _maxabs(params["itm_head.0.weight"].grad, 2 * g1) <= 2e-2 * float(g1.abs().max())
    for p in params.values():
        p.grad = None
    out3 = _run(model, g, batch)
    (out3["loss_itc"] + out3["loss_itm"] + out3["loss_mlm"] + out3["loss_mim"]).backward()
    assert _maxabs(params["itm_head.0.weight"].grad, g1) <= 2e-2 * float(g1.abs().max())


@pytest.mark.parametrize("name", ["base_mse.pt", "base_vq.pt"])
def test_base_config_against_reference(golden_dir, name):
    g = _load(golden_dir, name)
    model, cfg = _build(g)
    batch = _batch(g, cfg)
    model._vis.collect, model._txt.collect, model._fus.collect = [], [], []
    with torch.no_grad():
        out = _run(model, g, batch)
    tok, nd, B = [0, 1, 7, -1], 16, g["B"]
    nv, nt, nf = cfg["vision_depth"], cfg["text_layers"], cfg["fusion_layers"]
    groups = {"vision": [a[:B] for a in model._vis.collect[:nv]], "vision_masked": [a[B:] for a in model._vis.collect[:nv]],
              "text": [a[:B] for a in model._txt.collect[:nt]], "text_masked": [a[B:] for a in model._txt.collect[:nt]],
              "fusion_pos": [a[:B] for a in model._fus.collect[:nf]]}
    worst = 0.0
    for grp, acts in groups.items():
        for i, (a, s) in enumerate(zip(acts, g["acts_summary"][grp])):
            worst = max(worst, _maxabs(a[:, tok, :nd], s["slice"]))
    assert worst <= 2e-2, worst
    for k, v in g["losses"].items():
        tol = 5e-3 if (k == "loss_mim" and cfg["use_vision_tokenizer"]) else 1e-3
        assert abs(float(out[k]) - v) <= tol * max(1.0, abs(v)), (k, float(out[k]), v)


def test_train_mode_runs_and_dropout_changes_losses(golden_dir):
    g = _load(golden_dir, "tiny_mse.pt")
    model, cfg = _build(g, train=True)
    batch = _batch(g, cfg)
    model._forced_masks = g["ids_mask"]
    out = model(batch["image"], batch["text_ids"], batch["text_atts"], text_ids_masked=batch["text_ids_masked"],
                masked_pos=batch["masked_pos"], masked_ids=batch["masked_ids"], ret_mim_loss=True, data_source="image")
    total = out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + out["loss_mim"]
    total.backward()
    assert torch.isfinite(total)
    gsum = sum(float(p.grad.abs().sum()) for p in model.parameters() if p.grad is not None)
    assert gsum > 0 and np.isfinite(gsum)
    assert abs(float(out["loss_mlm"]) - g["losses"]["loss_mlm"]) > 1e-6  # dropout active


def test_build_mlp_head_matches_torch():
    """build_mlp (xfm.py:115-121), as used by model_nlvr.py:25 on the concatenated CLS rows: forward and all gradients
    against the same nn.Sequential evaluated by torch in fp32."""
    from xfm_b200.xfm import build_mlp
    torch.manual_seed(3)
    head = build_mlp(256, 2).cuda()
    ref = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.LayerNorm(512), torch.nn.GELU(), torch.nn.Linear(512, 2))
    ref.load_state_dict({k: v.detach().cpu() for k, v in head.state_dict().items()})
    x = torch.randn(24, 256)
    xr = x.clone().requires_grad_(True)
    xg = x.cuda().requires_grad_(True)
    tgt = torch.randint(0, 2, (24,))
    lr = torch.nn.functional.cross_entropy(ref(xr), tgt)
    lg = torch.nn.functional.cross_entropy(head(xg), tgt.cuda())
    lr.backward()
    lg.backward()
    assert abs(float(lg) - float(lr)) < 5e-3
    assert float((xg.grad.cpu() - xr.grad).abs().max()) < 2e-2 * float(xr.grad.abs().max()) + 1e-5
    for (n, p), (_, q) in zip(head.named_parameters(), ref.named_parameters()):
        assert float((p.grad.cpu() - q.grad).abs().max()) < 3e-2 * float(q.grad.abs().max()) + 1e-5, n


def test_fused_itm_mlm_pass_equals_separate_calls(golden_dir):
    """model_pretrain's single 4B-sample fusion pass (get_matching_and_fuse_mlm_loss) against the reference call pattern
    get_matching_loss + get_fuse_mlm_loss (xfm.py:749-802, 638-656): same losses, same parameter gradients."""
    g = _load(golden_dir, "tiny_vq.pt")
    res = {}
    for fused in (True, False):
        model, cfg = _build(g)
        model.fuse_itm_mlm = fused
        batch = _batch(g, cfg)
        out = _run(model, g, batch)
        (out["loss_itm"] + out["loss_mlm"]).backward()
        res[fused] = (float(out["loss_itm"]), float(out["loss_mlm"]),
                      {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    a, b = res[True], res[False]
    assert abs(a[0] - b[0]) < 1e-5 and abs(a[1] - b[1]) < 1e-5, (a[:2], b[:2])
    assert set(a[2]) == set(b[2])
    for n in a[2]:
        scale = max(float(b[2][n].abs().max()), 1e-8)
        # bf16 gradient tensors summed in a different grouping; key biases have a theoretically zero gradient (pure noise)
        assert _maxabs(a[2][n], b[2][n]) <= 2e-2 * scale + 1e-5, n


def _grad_check(model, ref_grads, names, tol=6e-2):
    params = dict(model.named_parameters())
    for n in names:
        mine, ref = params[n].grad, ref_grads[n]
        assert mine is not None, n
        assert _maxabs(mine, ref) <= tol * max(float(ref.abs().max()), 1e-8), n


def test_text_only_mlm_stream_against_oracle(record):
    """XFMBase.get_mlm_loss / XFM.forward_text (xfm.py:805-812, model_pretrain.py:93-98): the text-only stream — text encoder
    on the masked ids, LM head on the gathered positions, CE(ignore -100) — loss and gradients, no image involved."""
    from xfm_b200.model_pretrain import XFM
    cfg = O.tiny_config()
    B, Lt, Mm = 5, 24, 6
    sd = O.make_state_dict(cfg, 0)
    for v in sd.values():
        if v.dtype.is_floating_point:
            v.requires_grad_(True)
    batch = O.make_batch(cfg, B, L=Lt, M=Mm, seed=21)
    ref = O.text_mlm_loss(batch["text_ids_masked"], batch["text_atts"], batch["masked_pos"], batch["masked_ids"], sd, cfg)
    ref.backward()
    model = XFM(dict(cfg), init=lambda n, s: O.make_tensor(n, s, 0), device="cuda").eval()
    b = {k: v.cuda() for k, v in batch.items()}
    out = model(None, b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
                masked_ids=b["masked_ids"])
    assert set(out) == {"loss_mlm"}
    rel = abs(float(out["loss_mlm"]) - float(ref)) / max(1.0, abs(float(ref)))
    record("text_mlm", mine=float(out["loss_mlm"]), oracle=float(ref), rel=rel)
    assert rel <= 1e-3, (float(out["loss_mlm"]), float(ref))
    direct = model.get_mlm_loss(b["text_ids_masked"], b["text_atts"], None, None, b["masked_pos"], b["masked_ids"])
    assert abs(float(direct) - float(out["loss_mlm"])) < 1e-6
    out["loss_mlm"].backward()
    grads = {k: v.grad for k, v in sd.items() if v.dtype.is_floating_point and v.grad is not None}
    _grad_check(model, grads, ["text_encoder.lm_head.dense.weight", "text_encoder.lm_head.layer_norm.weight",
                               "text_encoder.lm_head.bias", "text_encoder.roberta.embeddings.word_embeddings.weight",
                               "text_encoder.roberta.encoder.layer.1.attention.self.query.weight",
                               "text_encoder.roberta.encoder.layer.0.output.dense.weight",
                               "text_encoder.roberta.embeddings.position_embeddings.weight"])
    params = dict(model.named_parameters())
    for n in ("vision_encoder.blocks.0.mlp.fc1.weight", "fusion_encoder.roberta.encoder.layer.0.output.dense.weight",
              "itm_head.0.weight"):
        assert params[n].grad is None, n     # nothing outside the text encoder takes part (xfm.py:805-812)


def test_region_branch_against_reference(golden_dir, record):
    """the region / bbox branch of the pre-training forward (model_pretrain.py:39-41,81-86) against the fixture
    the UNMODIFIED reference produced — idx_to_group_img gather + region-weighted pooling (beit2.py:468-475), region masks as
    cross-attention key masks, predict_bbox, L1 + GIoU (xfm.py:815-854) — losses, embeddings, boxes and gradients."""
    from xfm_b200.model_pretrain import XFM
    g = _load(golden_dir, "tiny_region.pt")
    cfg = g["cfg"]
    model = XFM(dict(cfg), init=lambda n, s: O.make_tensor(n, s, 0), device="cuda").eval()
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.make_region_batch(cfg).items()}
    with torch.no_grad():
        emb, atts, full = model.get_vision_embeds(b["image"], image_atts=b["image_atts"], idx_to_group_img=b["idx_to_group_img"])
        only_full, ones = model.get_vision_embeds(b["image"], idx_to_group_img=b["idx_to_group_img"])
    assert atts is b["image_atts"] and bool((ones == 1).all())
    assert _maxabs(emb, g["image_embeds"]) <= 2e-2 and _maxabs(full, g["image_embeds_fullatts"]) <= 2e-2
    assert _maxabs(only_full, g["image_embeds_fullatts"]) <= 2e-2
    for fuse in (True, False):
        model.zero_grad()
        model.fuse_itm_mlm = fuse
        model._forced_negatives = (g["image_neg_idx"], g["text_neg_idx"])
        out = model(b["image"], b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
                    masked_ids=b["masked_ids"], image_atts=b["image_atts"], idx_to_group_img=b["idx_to_group_img"],
                    target_bbox=b["target_bbox"], is_image=b["is_image"], ret_mim_loss=True, ret_bbox_loss=True,
                    ret_bbox_giou=True, data_source="region")
        assert float(out["loss_mim"]) == 0.0
        for k in ("loss_itc", "loss_itm", "loss_mlm", "loss_bbox", "loss_giou"):
            rel = abs(float(out[k]) - g["losses"][k]) / max(1.0, abs(g["losses"][k]))
            record("region_loss", fused=fuse, loss=k, mine=float(out[k]), reference=g["losses"][k], rel=rel)
            assert rel <= (2e-3 if k == "loss_itc" else 1e-3), (fuse, k, float(out[k]), g["losses"][k])
        total = out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + out["loss_bbox"] + out["loss_giou"]
        total.backward()
        _grad_check(model, g["grads"], list(g["grads"]))


def test_bert_text_encoder_variant_against_reference(golden_dir, record):
    """SURVEY §8 row x2 (north_star names models/xbert.py BertSelfAttention): a BERT-class text encoder — BertForMaskedLM
    parameter names, absolute position ids, padding_idx 0, LayerNorm eps 1e-12 — on the same kernels, against the fixture
    the UNMODIFIED reference produced (identical for both placements of 1/sqrt(d), xbert.py:296-301,329-330)."""
    from xfm_b200.model_pretrain import XFM
    g = _load(golden_dir, "tiny_bert.pt")
    cfg = g["cfg"]
    model = XFM(dict(cfg), init=lambda n, s: O.make_tensor(n, s, 0), device="cuda").eval()
    sd_names = set(model.state_dict())
    for n in ("text_encoder.bert.embeddings.position_ids", "text_encoder.cls.predictions.decoder.weight",
              "text_encoder.cls.predictions.transform.LayerNorm.weight", "fusion_encoder.roberta.embeddings.word_embeddings.weight"):
        assert n in sd_names, n
    assert not any(k.startswith("text_encoder.roberta.") or k.startswith("text_encoder.lm_") for k in sd_names)
    b = {k: v.cuda() for k, v in O.make_batch(cfg, 4, L=24, M=6, seed=1).items()}
    with torch.no_grad():
        te = model.get_text_embeds(b["text_ids"], b["text_atts"])
    assert _maxabs(te, g["text_embeds"]) <= 2e-2
    model._forced_negatives = (g["image_neg_idx"], g["text_neg_idx"])
    out = model(b["image"], b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
                masked_ids=b["masked_ids"], ret_mim_loss=False, data_source="image")
    for k, v in g["losses"].items():
        rel = abs(float(out[k]) - v) / max(1.0, abs(v))
        record("bert_variant", loss=k, mine=float(out[k]), reference=v, rel=rel)
        assert rel <= (2e-3 if k == "loss_itc" else 1e-3), (k, float(out[k]), v)
    (out["loss_itc"] + out["loss_itm"] + out["loss_mlm"]).backward()
    _grad_check(model, g["grads"], list(g["grads"]))
    model.zero_grad()
    t = model(None, b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
              masked_ids=b["masked_ids"])["loss_mlm"]
    assert abs(float(t) - g["text_only_mlm"]) <= 1e-3 * abs(g["text_only_mlm"])


def test_retrieval_model_against_oracle():
    """models/model_retrieval.py:26-37 (BASELINE config #3 shape, tiny widths): ITC with idx soft labels + idx-masked
    hard-negative ITM, text gradients through the fusion encoder (is_pretrain=False)."""
    from xfm_b200.model_retrieval import XFMForRetrieval
    cfg = O.tiny_config()
    B, Lt = 6, 24
    sd = O.make_state_dict(cfg, 0)
    for v in sd.values():
        if v.dtype.is_floating_point:
            v.requires_grad_(True)
    batch = O.make_batch(cfg, B, L=Lt, M=6, seed=3)
    idx = torch.tensor([0, 1, 0, 2, 1, 3])
    ie = O.vision_forward(batch["image"], sd, cfg)
    ia = torch.ones(ie.shape[:2], dtype=torch.long)
    te = O.text_forward(batch["text_ids"], batch["text_atts"], sd, cfg)
    fi, ft = O.get_features(ie, te, sd)
    w_i2t, w_t2i = O.hard_negative_weights(fi.detach(), ft.detach(), sd["temp"].detach(), idx)
    tneg, ineg = w_i2t.argmax(1), w_t2i.argmax(1)   # deterministic stand-in for torch.multinomial (xfm.py:736-746)
    itc = O.contrastive_loss(fi, ft, sd["temp"], idx)
    itm, _ = O.matching_loss(ie, ia, te, batch["text_atts"], ineg, tneg, sd, cfg, is_pretrain=False)
    (itc + itm).backward()
    model = XFMForRetrieval(dict(cfg), init=lambda n, s: O.make_tensor(n, s, 0), device="cuda").eval()
    model._forced_negatives = (ineg, tneg)
    b = {k: v.cuda() for k, v in batch.items()}
    l_itc, l_itm = model(b["image"], b["text_ids"], b["text_atts"], idx=idx.cuda())
    assert abs(float(l_itc) - float(itc)) <= 2e-3 * max(1.0, abs(float(itc))), (float(l_itc), float(itc))
    assert abs(float(l_itm) - float(itm)) <= 1e-3 * max(1.0, abs(float(itm))), (float(l_itm), float(itm))
    (l_itc + l_itm).backward()
    ref = {k: v.grad for k, v in sd.items() if v.dtype.is_floating_point and v.grad is not None}
    _grad_check(model, ref, ["itm_head.0.weight", "fusion_encoder.roberta.encoder.layer.1.crossattention.self.key.weight",
                             "text_encoder.roberta.encoder.layer.0.intermediate.dense.weight", "vision_proj.weight",
                             "vision_encoder.blocks.1.mlp.fc2.weight"])


def test_nlvr_model_against_oracle():
    """models/model_nlvr.py:28-44 (BASELINE config #4 shape, tiny widths): two images per text, two fusion passes sharing
    the text, concatenated CLS -> build_mlp -> CE."""
    from xfm_b200.model_nlvr import XFMForNLVR
    cfg = O.tiny_config()
    B, Lt = 4, 24
    sd = O.make_state_dict(cfg, 0)
    batch = O.make_batch(cfg, 2 * B, L=Lt, M=6, seed=5)
    image = batch["image"]
    text_ids, text_atts = batch["text_ids"][:B], batch["text_atts"][:B]
    targets = torch.tensor([0, 1, 1, 0])
    model = XFMForNLVR(dict(cfg), init=lambda n, s: O.make_tensor(n, s, 0), device="cuda").eval()
    head = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.cls_head.state_dict().items()}
    ie = O.vision_forward(image, sd, cfg)
    ia = torch.ones(ie.shape[:2], dtype=torch.long)
    te = O.text_forward(text_ids, text_atts, sd, cfg)
    c1 = O.fusion_forward(te, text_atts, ie[:B], ia[:B], sd, cfg)[:, 0]
    c2 = O.fusion_forward(te, text_atts, ie[B:], ia[B:], sd, cfg)[:, 0]
    x = torch.cat([c1, c2], -1)
    x = torch.nn.functional.linear(x, head["0.weight"], head["0.bias"])
    x = torch.nn.functional.gelu(torch.nn.functional.layer_norm(x, (x.shape[-1],), head["1.weight"], head["1.bias"], 1e-5))
    pred = torch.nn.functional.linear(x, head["3.weight"], head["3.bias"])
    ref_loss = torch.nn.functional.cross_entropy(pred, targets)
    ref_loss.backward()
    loss = model(image.cuda(), text_ids.cuda(), text_atts.cuda(), targets.cuda())
    assert abs(float(loss) - float(ref_loss)) <= 2e-3 * max(1.0, abs(float(ref_loss))), (float(loss), float(ref_loss))
    loss.backward()
    for n, p in model.cls_head.named_parameters():
        assert _maxabs(p.grad, head[n].grad) <= 6e-2 * max(float(head[n].grad.abs().max()), 1e-8), n
    pred_gpu = model(image.cuda(), text_ids.cuda(), text_atts.cuda(), targets.cuda(), train=False)
    assert _maxabs(pred_gpu, pred.detach()) <= 2e-2


def test_retrieval_evaluation_rerank_against_oracle():
    """Retrieval.py:76-184: ITC top-k candidates re-scored by the fusion encoder + itm_head, both
    directions.  Oracle: the reference's per-row loop restated on the CPU functions (every (image, text) pair scored, so the
    comparison does not depend on which near-tied candidates a side picks).  Also the rank sharding of the loop."""
    from xfm_b200.model_retrieval import XFMForRetrieval
    from xfm_b200 import retrieval_eval as RE
    cfg = O.tiny_config()
    n_img, n_txt, Lt, k = 5, 11, 24, 3
    sd = O.make_state_dict(cfg, 0)
    bi = O.make_batch(cfg, n_img, L=Lt, M=6, seed=11)
    bt = O.make_batch(cfg, n_txt, L=Lt, M=6, seed=12)
    with torch.no_grad():
        ie = O.vision_forward(bi["image"], sd, cfg)
        te = O.text_forward(bt["text_ids"], bt["text_atts"], sd, cfg)
        fi, ft = O.get_features(ie, te, sd)
        sims = fi @ ft.t()
        full = torch.empty(n_img, n_txt)
        for i in range(n_img):   # Retrieval.py:139-147 with every text as a candidate
            enc = ie[i].repeat(n_txt, 1, 1)
            out = O.fusion_forward(te, bt["text_atts"], enc, torch.ones(enc.shape[:2], dtype=torch.long), sd, cfg)
            full[i] = O.itm_head(out[:, 0, :], sd)[:, 1]
    model = XFMForRetrieval(dict(cfg), init=lambda n, s: O.make_tensor(n, s, 0), device="cuda").eval()
    s_i2t, s_t2i = RE.evaluation(model, [bi["image"][:2].cuda(), bi["image"][2:].cuda()], bt["text_ids"].cuda(),
                                 bt["text_atts"].cuda(), {"k_test": k, "batch_size_test_text": 4})
    assert s_i2t.shape == (n_img, n_txt) and s_t2i.shape == (n_txt, n_img)
    tol = 2e-2 * max(1.0, float(full.abs().max()))
    for mat, ref, ref_sims in ((s_i2t, full, sims), (s_t2i, full.t(), sims.t())):
        mat = torch.from_numpy(mat)
        scored = mat != -100.0
        assert (scored.sum(1) == k).all()
        assert float((mat - ref)[scored].abs().max()) <= tol
        # the scored candidates are the ITC top-k (ties at the k-th similarity may resolve either way)
        kth = ref_sims.topk(k, dim=1).values[:, -1:]
        assert bool((ref_sims[scored].view(-1, k) >= kth - 2e-3).all())
    # rank sharding (Retrieval.py:133-136,153-155): the two ranks' matrices sum to the single-rank result - 100
    t32, t16, temb = RE.encode_texts(model, bt["text_ids"].cuda(), bt["text_atts"].cuda())
    i16, iemb = RE.encode_images(model, [bi["image"].cuda()])
    parts = [RE.rerank(model, i16, iemb, t32, t16, temb, bt["text_atts"].cuda(), k, pairs_per_pass=7, shard=(r, 2)) for r in (0, 1)]
    one = RE.rerank(model, i16, iemb, t32, t16, temb, bt["text_atts"].cuda(), k)
    for d in (0, 1):
        assert float((parts[0][d] + parts[1][d] - (one[d] - 100.0)).abs().max()) <= 2e-2



def _vqa_inputs(b):
    import types
    c = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    q = types.SimpleNamespace(input_ids=c["q_ids"], attention_mask=c["q_atts"])
    a = types.SimpleNamespace(input_ids=c["a_ids"], attention_mask=c["a_atts"])
    cand = types.SimpleNamespace(input_ids=c["cand_ids"], attention_mask=c["cand_atts"])
    return c, q, a, cand


def test_vqa_model_against_reference(golden_dir):
    """models/model_generation.py:93-202 (BASELINE config #5 shape, tiny widths) against the fixture produced by the
    UNMODIFIED reference: question states, per-answer causal-decoder losses, the weighted loss, parameter gradients through
    decoder / fusion / text / vision, and the rank_answer re-ranking."""
    from xfm_b200.model_generation import XFMForVQA
    g = _load(golden_dir, "tiny_vqa.pt")
    cfg = g["cfg"]
    b = O.make_vqa_batch(cfg)
    model = XFMForVQA(dict(cfg, num_dec_layers=cfg["dec_layers"], decoder_fusion_start_at=0, pad_token_id=cfg["pad_id"]),
                      init=lambda n, s: O.make_tensor(n, s, 0), device="cuda").eval()
    c, q, a, cand = _vqa_inputs(b)
    with torch.no_grad():
        ie, ia = model.get_vision_embeds(c["image"])
        te = model.get_text_embeds(q.input_ids, q.attention_mask)
        qo = model.get_cross_embeds(ie, ia, text_embeds=te, text_atts=q.attention_mask, is_pretrain=False)
    assert _maxabs(qo, g["question_output"]) <= 2e-2
    loss = model(c["image"], q, a, k=b["k"], weights=c["weights"], train=True)
    assert _maxabs(model.last_answer_loss, g["answer_loss"]) <= 1e-3 * float(g["answer_loss"].abs().max())
    assert abs(float(loss) - g["loss"]) <= 1e-3 * abs(g["loss"]), (float(loss), g["loss"])
    loss.backward()
    _grad_check(model, g["grads"], list(g["grads"]))
    ids, probs = model(c["image"], q, cand, k=g["k_test"], train=False)
    assert torch.equal(ids.cpu(), g["topk_ids"])
    torch.testing.assert_close(probs.cpu(), g["topk_probs"], rtol=0.1, atol=1e-4)
    # state_dict carries the reference's decoder names, incl. the tied head
    sd = model.state_dict()
    for n in ("text_decoder.roberta.encoder.layer.1.crossattention.self.key.weight", "text_decoder.lm_head.decoder.weight",
              "text_decoder.lm_head.decoder.bias", "text_decoder.roberta.embeddings.position_ids"):
        assert n in sd, n
    assert sd["text_decoder.lm_head.decoder.weight"].data_ptr() == sd["text_decoder.roberta.embeddings.word_embeddings.weight"].data_ptr()


def test_vqa_base_width_against_oracle():
    """Same path at the XFM-base widths (768 wide, vocabulary 50265, 40-token questions, 8-token answers; 2 layers per
    stack so the CPU oracle finishes in seconds), train mode off: loss, gradients, ranking."""
    from xfm_b200.model_generation import XFMForVQA
    cfg = O.base_config(vision_depth=2, text_layers=2, fusion_layers=2, dec_layers=2, use_bbox=False)
    sd = O.make_state_dict(cfg, 0)
    for v in sd.values():
        if v.dtype.is_floating_point:
            v.requires_grad_(True)
    b = O.make_vqa_batch(cfg, B=4, L=40, La=8, n_cand=9, seed=17)
    col = {}
    ref = O.vqa_train_loss(b["image"], b["q_ids"], b["q_atts"], b["a_ids"], b["a_atts"], b["k"], b["weights"], sd, cfg,
                           collect=col)
    ref.backward()
    with torch.no_grad():
        ref_ids, ref_probs = O.vqa_rank(b["image"], b["q_ids"], b["q_atts"], b["cand_ids"], b["cand_atts"], 4, sd, cfg)
    model = XFMForVQA(dict(cfg, num_dec_layers=2, pad_token_id=cfg["pad_id"]), init=lambda n, s: O.make_tensor(n, s, 0),
                      device="cuda").eval()
    c, q, a, cand = _vqa_inputs(b)
    loss = model(c["image"], q, a, k=b["k"], weights=c["weights"], train=True)
    assert _maxabs(model.last_answer_loss, col["answer_loss"].detach()) <= 2e-3 * float(col["answer_loss"].abs().max())
    assert abs(float(loss) - float(ref)) <= 1e-3 * abs(float(ref)), (float(loss), float(ref))
    loss.backward()
    grads = {k: v.grad for k, v in sd.items() if v.dtype.is_floating_point and v.grad is not None}
    _grad_check(model, grads, ["text_decoder.roberta.encoder.layer.0.crossattention.self.value.weight",
                               "text_decoder.roberta.encoder.layer.1.attention.self.key.weight",
                               "text_decoder.lm_head.layer_norm.weight",
                               "fusion_encoder.roberta.encoder.layer.1.output.dense.weight",
                               "text_encoder.roberta.encoder.layer.0.attention.self.query.weight",
                               "vision_encoder.blocks.0.attn.qkv.weight"])
    ids, probs = model(c["image"], q, cand, k=4, train=False)
    # candidates whose probabilities are within bf16 noise of each other may swap places: compare as sets + sorted values
    assert torch.equal(ids.cpu().sort(1).values, ref_ids.sort(1).values)
    # 50265-way random-init logits make the re-ranked probabilities span 27 decades: compare them in log space
    torch.testing.assert_close(probs.cpu().double().log(), ref_probs.double().log(), rtol=0, atol=0.3)
    model.train()
    l2 = model(c["image"], q, a, k=b["k"], weights=c["weights"], train=True)
    l2.backward()
    assert torch.isfinite(l2) and abs(float(l2) - float(loss)) > 0


def test_nlvr_base_width_head_against_torch():
    """model_nlvr.py:25 at the XFM-base width: build_mlp(1536 -> 3072 -> 2) needs a 3072-wide LayerNorm (beyond the
    2048-wide rows of every encoder LayerNorm).  2 layers per stack; loss and head gradients against torch on the module's
    own concatenated CLS rows."""
    from xfm_b200.model_nlvr import XFMForNLVR
    cfg = O.base_config(vision_depth=2, text_layers=2, fusion_layers=2, use_bbox=False)
    B = 4
    batch = O.make_batch(cfg, 2 * B, L=40, M=2, seed=9)
    model = XFMForNLVR(dict(cfg), init=lambda n, s: O.make_tensor(n, s, 0), device="cuda").eval()
    image, ids, atts = batch["image"].cuda(), batch["text_ids"][:B].cuda(), batch["text_atts"][:B].cuda()
    targets = torch.tensor([0, 1, 1, 0], device="cuda")
    loss = model(image, ids, atts, targets)
    loss.backward()
    with torch.no_grad():
        ie, ia = model.get_vision_embeds(image)
        te = model.get_text_embeds(ids, atts)
        c1 = model.get_cross_embeds(ie[:B], ia[:B], text_embeds=te, text_atts=atts, is_pretrain=False)[:, 0]
        c2 = model.get_cross_embeds(ie[B:], ia[B:], text_embeds=te, text_atts=atts, is_pretrain=False)[:, 0]
    x = torch.cat([c1, c2], -1).float().cpu()
    head = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.cls_head.state_dict().items()}
    h = torch.nn.functional.linear(x, head["0.weight"], head["0.bias"])
    h = torch.nn.functional.gelu(torch.nn.functional.layer_norm(h, (h.shape[-1],), head["1.weight"], head["1.bias"], 1e-5))
    ref = torch.nn.functional.cross_entropy(torch.nn.functional.linear(h, head["3.weight"], head["3.bias"]), targets.cpu())
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 2e-3 * max(1.0, abs(float(ref))), (float(loss), float(ref))
    for n, p in model.cls_head.named_parameters():
        assert _maxabs(p.grad, head[n].grad) <= 6e-2 * max(float(head[n].grad.abs().max()), 1e-8), n
