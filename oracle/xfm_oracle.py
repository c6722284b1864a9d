"""CPU oracle for the XFM pre-training / fine-tuning hot path.

TEST INFRASTRUCTURE ONLY.  This is a plain-PyTorch fp32 *restatement* of the reference
(zhangxinsong-nlp/XFM) arithmetic, written functionally over a state_dict with the reference's
parameter names.  It is the checker the CUDA path is compared against in
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg; nothing
under `xfm_b200/` imports it and the product path never routes through it.

Parity pin: the reference ships no tests or golden vectors.  This oracle is
pinned against the reference's own modules, imported unmodified in the build container through
`oracle/ref_shim.py`; `tools/make_golden.py` generated `tests/golden/*.pt` from the REFERENCE and
`tests/test_oracle_golden.py` checks this file against those fixtures.

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
import math
import random as _pyrandom
import zlib

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------------------


def base_config(**over):
    """XFM-base dimensions (configs/xfm-pt/Pretrain_XBrain_base_4m.yaml:64-75, roberta-base config,
    models/beit2.py:540-545)."""
    cfg = dict(
        image_res=224, patch_size=16, vision_width=768, vision_depth=12, vision_heads=12, vision_mlp=3072,
        vision_ln_eps=1e-6, init_values=0.1, drop_path_rate=0.1,
        vocab_size=50265, hidden=768, text_layers=12, fusion_layers=12, heads=12, ffn=3072, max_pos=514,
        type_vocab=1, pad_id=1, ln_eps=1e-5, hidden_dropout=0.1, attn_dropout=0.1,
        embed_dim=256, temp=0.07, min_temp=0.001, max_temp=0.5,
        num_masking_patches=75, min_num_patches=16,
        use_vision_tokenizer=False, codebook_size=8192, codebook_dim=32,
        use_bbox=True,
        text_arch="roberta",   # "bert": the models/xbert.py text encoder (BertForMaskedLM naming, absolute position ids)
    )
    cfg.update(over)
    return cfg


def tiny_config(**over):
    """Small configuration with the same structure (head_dim stays 64) for fast CPU/GPU parity tests."""
    cfg = base_config(
        image_res=64, vision_width=128, vision_depth=2, vision_heads=2, vision_mlp=512,
        vocab_size=1000, hidden=128, text_layers=2, fusion_layers=2, heads=2, ffn=512, max_pos=66,
        embed_dim=64, num_masking_patches=6, min_num_patches=2, codebook_size=512, codebook_dim=32)
    cfg.update(over)
    return cfg


def num_patches(cfg):
    return (cfg["image_res"] // cfg["patch_size"]) ** 2


# ----------------------------------------------------------------------------------------------
# deterministic synthetic weights (shared by the oracle, the reference shim and the CUDA path)
# ----------------------------------------------------------------------------------------------


def relative_position_index(ws):
    """models/beit2.py:92-109 — pair-wise relative position index incl. the 3 cls entries."""
    num_rel = (2 * ws - 1) * (2 * ws - 1) + 3
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = cf[:, :, None] - cf[:, None, :]
    rel = rel.permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    idx = torch.zeros((ws * ws + 1,) * 2, dtype=rel.dtype)
    idx[1:, 1:] = rel.sum(-1)
    idx[0, 0:] = num_rel - 3
    idx[0:, 0] = num_rel - 2
    idx[0, 0] = num_rel - 1
    return idx


def param_shapes(cfg):
    """Parameter / buffer names and shapes of the pre-train module."""
    D, Fv, Hv = cfg["vision_width"], cfg["vision_mlp"], cfg["vision_heads"]
    ws = cfg["image_res"] // cfg["patch_size"]
    P = cfg["patch_size"]
    H, Ff, V = cfg["hidden"], cfg["ffn"], cfg["vocab_size"]
    s = {}
    s["temp"] = ()
    v = "vision_encoder."
    s[v + "cls_token"] = (1, 1, D)
    s[v + "mask_token"] = (1, 1, D)
    s[v + "patch_embed.proj.weight"] = (D, 3, P, P)
    s[v + "patch_embed.proj.bias"] = (D,)
    for i in range(cfg["vision_depth"]):
        b = f"{v}blocks.{i}."
        s[b + "gamma_1"] = (D,)
        s[b + "gamma_2"] = (D,)
        s[b + "norm1.weight"] = (D,)
        s[b + "norm1.bias"] = (D,)
        s[b + "attn.q_bias"] = (D,)
        s[b + "attn.v_bias"] = (D,)
        s[b + "attn.relative_position_bias_table"] = ((2 * ws - 1) ** 2 + 3, Hv)
        s[b + "attn.qkv.weight"] = (3 * D, D)
        s[b + "attn.proj.weight"] = (D, D)
        s[b + "attn.proj.bias"] = (D,)
        s[b + "norm2.weight"] = (D,)
        s[b + "norm2.bias"] = (D,)
        s[b + "mlp.fc1.weight"] = (Fv, D)
        s[b + "mlp.fc1.bias"] = (Fv,)
        s[b + "mlp.fc2.weight"] = (D, Fv)
        s[b + "mlp.fc2.bias"] = (D,)
    s[v + "fc_norm.weight"] = (D,)
    s[v + "fc_norm.bias"] = (D,)

    def roberta(prefix, layers, cross, kin=D, heads=("lm_head", "lm_cap_head"), arch="roberta"):
        if arch == "bert":
            return bert(prefix, layers)
        e = prefix + "roberta.embeddings."
        s[e + "word_embeddings.weight"] = (V, H)
        s[e + "position_embeddings.weight"] = (cfg["max_pos"], H)
        s[e + "token_type_embeddings.weight"] = (cfg["type_vocab"], H)
        s[e + "LayerNorm.weight"] = (H,)
        s[e + "LayerNorm.bias"] = (H,)
        for i in range(layers):
            l = f"{prefix}roberta.encoder.layer.{i}."
            for att, kin in (("attention", H),) + ((("crossattention", kin),) if cross else ()):
                s[l + att + ".self.query.weight"] = (H, H)
                s[l + att + ".self.query.bias"] = (H,)
                s[l + att + ".self.key.weight"] = (H, kin)
                s[l + att + ".self.key.bias"] = (H,)
                s[l + att + ".self.value.weight"] = (H, kin)
                s[l + att + ".self.value.bias"] = (H,)
                s[l + att + ".output.dense.weight"] = (H, H)
                s[l + att + ".output.dense.bias"] = (H,)
                s[l + att + ".output.LayerNorm.weight"] = (H,)
                s[l + att + ".output.LayerNorm.bias"] = (H,)
            s[l + "intermediate.dense.weight"] = (Ff, H)
            s[l + "intermediate.dense.bias"] = (Ff,)
            s[l + "output.dense.weight"] = (H, Ff)
            s[l + "output.dense.bias"] = (H,)
            s[l + "output.LayerNorm.weight"] = (H,)
            s[l + "output.LayerNorm.bias"] = (H,)
        for head in heads:
            h = prefix + head + "."
            s[h + "bias"] = (V,)
            s[h + "dense.weight"] = (H, H)
            s[h + "dense.bias"] = (H,)
            s[h + "layer_norm.weight"] = (H,)
            s[h + "layer_norm.bias"] = (H,)
            if head == "lm_cap_head":  # untied (SURVEY.md §7 "Weight tying")
                s[h + "decoder.weight"] = (V, H)

    def bert(prefix, layers):
        """BertForMaskedLM without cross-attention (models/xbert.py:1523-1540, 663-707)."""
        e = prefix + "bert.embeddings."
        s[e + "word_embeddings.weight"] = (V, H)
        s[e + "position_embeddings.weight"] = (cfg["max_pos"], H)
        s[e + "token_type_embeddings.weight"] = (cfg["type_vocab"], H)
        s[e + "LayerNorm.weight"] = (H,)
        s[e + "LayerNorm.bias"] = (H,)
        for i in range(layers):
            l = f"{prefix}bert.encoder.layer.{i}."
            for n, shape in (("attention.self.query.weight", (H, H)), ("attention.self.query.bias", (H,)),
                             ("attention.self.key.weight", (H, H)), ("attention.self.key.bias", (H,)),
                             ("attention.self.value.weight", (H, H)), ("attention.self.value.bias", (H,)),
                             ("attention.output.dense.weight", (H, H)), ("attention.output.dense.bias", (H,)),
                             ("attention.output.LayerNorm.weight", (H,)), ("attention.output.LayerNorm.bias", (H,)),
                             ("intermediate.dense.weight", (Ff, H)), ("intermediate.dense.bias", (Ff,)),
                             ("output.dense.weight", (H, Ff)), ("output.dense.bias", (H,)),
                             ("output.LayerNorm.weight", (H,)), ("output.LayerNorm.bias", (H,))):
                s[l + n] = shape
        h = prefix + "cls.predictions."
        s[h + "bias"] = (V,)
        s[h + "transform.dense.weight"] = (H, H)
        s[h + "transform.dense.bias"] = (H,)
        s[h + "transform.LayerNorm.weight"] = (H,)
        s[h + "transform.LayerNorm.bias"] = (H,)

    roberta("text_encoder.", cfg["text_layers"], cross=False, arch=cfg.get("text_arch", "roberta"))
    roberta("fusion_encoder.", cfg["fusion_layers"], cross=True)
    if cfg.get("dec_layers", 0) > 0:
        # XFMForVQA.text_decoder = RobertaForCausalLM (models/model_generation.py:41-54): cross-attention over the
        # question states (encoder_width = hidden) in every layer (decoder_fusion_start_at 0), one tied LM head
        roberta("text_decoder.", cfg["dec_layers"], cross=True, kin=H, heads=("lm_head",))
    E = cfg["embed_dim"]
    s["vision_proj.weight"] = (E, D)
    s["vision_proj.bias"] = (E,)
    s["text_proj.weight"] = (E, H)
    s["text_proj.bias"] = (E,)
    heads = [("itm_head", 2)] + ([("bbox_head", 4)] if cfg.get("use_bbox", True) else [])
    for name, nout in heads:
        s[f"{name}.0.weight"] = (2 * H, H)
        s[f"{name}.0.bias"] = (2 * H,)
        s[f"{name}.1.weight"] = (2 * H,)
        s[f"{name}.1.bias"] = (2 * H,)
        s[f"{name}.3.weight"] = (nout, 2 * H)
        s[f"{name}.3.bias"] = (nout,)
    if cfg["use_vision_tokenizer"]:
        s["lm_head.weight"] = (cfg["codebook_size"], D)
        s["lm_head.bias"] = (cfg["codebook_size"],)
        s.update(vqkd_param_shapes(cfg, "vqkd."))
    return s


def vqkd_param_shapes(cfg, prefix="vqkd."):
    """Encoder + task layer + codebook of the VQ-KD tokenizer (models/model_vqkd.py:242-246,293-309,
    models/vqkd_vit.py:285-318).  Decoder / teacher never run as a tokenizer and are not modelled."""
    D, Fv = 768 if cfg["vision_width"] == 768 else cfg["vision_width"], None
    Fv = 4 * D
    P = cfg["patch_size"]
    n = num_patches(cfg)
    s = {}
    e = prefix + "encoder."
    s[e + "cls_token"] = (1, 1, D)
    s[e + "pos_embed"] = (1, n + 1, D)
    s[e + "patch_embed.proj.weight"] = (D, 3, P, P)
    s[e + "patch_embed.proj.bias"] = (D,)
    for i in range(cfg["vision_depth"]):
        b = f"{e}blocks.{i}."
        s[b + "norm1.weight"] = (D,)
        s[b + "norm1.bias"] = (D,)
        s[b + "attn.q_bias"] = (D,)
        s[b + "attn.v_bias"] = (D,)
        s[b + "attn.qkv.weight"] = (3 * D, D)
        s[b + "attn.proj.weight"] = (D, D)
        s[b + "attn.proj.bias"] = (D,)
        s[b + "norm2.weight"] = (D,)
        s[b + "norm2.bias"] = (D,)
        s[b + "mlp.fc1.weight"] = (Fv, D)
        s[b + "mlp.fc1.bias"] = (Fv,)
        s[b + "mlp.fc2.weight"] = (D, Fv)
        s[b + "mlp.fc2.bias"] = (D,)
    s[e + "fc_norm.weight"] = (D,)
    s[e + "fc_norm.bias"] = (D,)
    s[prefix + "encode_task_layer.0.weight"] = (D, D)
    s[prefix + "encode_task_layer.0.bias"] = (D,)
    s[prefix + "encode_task_layer.2.weight"] = (cfg["codebook_dim"], D)
    s[prefix + "encode_task_layer.2.bias"] = (cfg["codebook_dim"],)
    s[prefix + "quantize.embedding.weight"] = (cfg["codebook_size"], cfg["codebook_dim"])
    return s


def _gen(name, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def make_tensor(name, shape, seed=0):
    """Deterministic synthetic value for one parameter, keyed by its NAME (order independent).

    Deliberately not the reference's init (zeros for biases, ones for LayerNorm): every parameter is
    non-trivial so a dropped bias / swapped gamma shows up in parity."""
    g = _gen(name, seed)
    leaf = name.rsplit(".", 1)[-1]
    if name == "temp":
        return torch.tensor(0.07)
    if leaf in ("gamma_1", "gamma_2"):
        return 0.1 + 0.02 * torch.randn(shape, generator=g)
    if "norm" in name.lower() and leaf == "weight" or name.endswith(".1.weight") and len(shape) == 1:
        return 1.0 + 0.05 * torch.randn(shape, generator=g)
    if leaf in ("bias", "q_bias", "v_bias"):
        return 0.02 * torch.randn(shape, generator=g)
    if leaf == "relative_position_bias_table":
        return 0.2 * torch.randn(shape, generator=g)
    if "quantize.embedding.weight" in name:
        return F.normalize(torch.randn(shape, generator=g), dim=-1)
    if "word_embeddings" in name or "position_embeddings" in name or "token_type" in name or "pos_embed" in name:
        return 0.05 * torch.randn(shape, generator=g)
    if "patch_embed.proj.weight" in name:
        return 0.03 * torch.randn(shape, generator=g)
    # dense weights: keep activations O(1) through depth
    fan_in = shape[-1] if len(shape) >= 2 else 1
    std = 0.6 / math.sqrt(max(fan_in, 1)) if len(shape) >= 2 else 0.02
    return std * torch.randn(shape, generator=g)


def make_state_dict(cfg, seed=0):
    sd = {k: make_tensor(k, shp, seed) for k, shp in param_shapes(cfg).items()}
    # weight tying: lm_head.decoder.weight IS word_embeddings.weight; decoder.bias IS bias
    return sd


def expand_tied(sd, cfg):
    """state_dict in the reference's full key layout (adds tied aliases and buffers)."""
    out = dict(sd)
    ws = cfg["image_res"] // cfg["patch_size"]
    rpi = relative_position_index(ws)
    for i in range(cfg["vision_depth"]):
        out[f"vision_encoder.blocks.{i}.attn.relative_position_index"] = rpi.clone()
    for p in ("text_encoder.", "fusion_encoder."):
        if p == "text_encoder." and cfg.get("text_arch", "roberta") == "bert":   # xbert.py:687-692,1536-1537
            out[p + "bert.embeddings.position_ids"] = torch.arange(cfg["max_pos"]).expand((1, -1)).clone()
            out[p + "cls.predictions.decoder.weight"] = out[p + "bert.embeddings.word_embeddings.weight"]
            out[p + "cls.predictions.decoder.bias"] = out[p + "cls.predictions.bias"]
            continue
        out[p + "roberta.embeddings.position_ids"] = torch.arange(cfg["max_pos"]).expand((1, -1)).clone()
        out[p + "lm_head.decoder.weight"] = out[p + "roberta.embeddings.word_embeddings.weight"]
        out[p + "lm_head.decoder.bias"] = out[p + "lm_head.bias"]
        out[p + "lm_cap_head.decoder.bias"] = out[p + "lm_cap_head.bias"]
    if cfg.get("dec_layers", 0) > 0:
        p = "text_decoder."
        out[p + "roberta.embeddings.position_ids"] = torch.arange(cfg["max_pos"]).expand((1, -1)).clone()
        out[p + "lm_head.decoder.weight"] = out[p + "roberta.embeddings.word_embeddings.weight"]
        out[p + "lm_head.decoder.bias"] = out[p + "lm_head.bias"]
    return out


# ----------------------------------------------------------------------------------------------
# synthetic batch
# ----------------------------------------------------------------------------------------------


def make_batch(cfg, B, L=40, M=15, seed=1, image_uniform=False):
    g = torch.Generator().manual_seed(seed)
    res = cfg["image_res"]
    if image_uniform:
        image = torch.rand(B, 3, res, res, generator=g)
    else:
        image = torch.randn(B, 3, res, res, generator=g)
    V = cfg["vocab_size"]
    text_ids = torch.randint(3, V - 1, (B, L), generator=g)
    text_ids[:, 0] = 0 if cfg["pad_id"] != 0 else 2   # bos / cls token: never the padding id
    text_atts = torch.ones(B, L, dtype=torch.long)
    for b in range(B):
        if b % 2 == 1:  # every other row padded to 3/4 length
            n_real = max(2, (3 * L) // 4)
            text_atts[b, n_real:] = 0
            text_ids[b, n_real:] = cfg["pad_id"]
    masked_pos = torch.zeros(B, M, dtype=torch.long)
    masked_ids = torch.full((B, M), -100, dtype=torch.long)
    text_ids_masked = text_ids.clone()
    for b in range(B):
        n_real = int(text_atts[b].sum())
        n_mask = min(M, max(1, (n_real - 1) // 2))
        perm = torch.randperm(n_real - 1, generator=g)[:n_mask] + 1
        perm, _ = torch.sort(perm)
        masked_pos[b, :n_mask] = perm
        masked_ids[b, :n_mask] = text_ids[b, perm]
        text_ids_masked[b, perm] = V - 1  # <mask>
    return dict(image=image, text_ids=text_ids, text_atts=text_atts, text_ids_masked=text_ids_masked,
                masked_pos=masked_pos, masked_ids=masked_ids)


# ----------------------------------------------------------------------------------------------
# MIM block-wise masking (host side, bit-exact contract)
# ----------------------------------------------------------------------------------------------


class MaskingGenerator:
    """models/masking_generator.py:26-105.  Consumes python `random` and `np.random` GLOBAL streams in
    the same order as the reference so that equal seeds give equal masks."""

    def __init__(self, input_size, num_masking_patches, min_num_patches=4, max_num_patches=None, min_aspect=0.3,
                 max_aspect=None):
        if not isinstance(input_size, tuple):
            input_size = (input_size,) * 2
        self.height, self.width = input_size
        self.num_masking_patches = num_masking_patches
        self.min_num_patches = min_num_patches
        self.max_num_patches = num_masking_patches if max_num_patches is None else max_num_patches
        max_aspect = max_aspect or 1 / min_aspect
        self.log_aspect_ratio = (math.log(min_aspect), math.log(max_aspect))

    def _mask(self, mask, max_mask_patches):  # :53-75
        delta = 0
        for _ in range(10):
            target_area = _pyrandom.uniform(self.min_num_patches, max_mask_patches)
            aspect_ratio = math.exp(_pyrandom.uniform(*self.log_aspect_ratio))
            h = int(round(math.sqrt(target_area * aspect_ratio)))
            w = int(round(math.sqrt(target_area / aspect_ratio)))
            if w < self.width and h < self.height:
                top = _pyrandom.randint(0, self.height - h)
                left = _pyrandom.randint(0, self.width - w)
                num_masked = mask[top:top + h, left:left + w].sum()
                if 0 < h * w - num_masked <= max_mask_patches:
                    for i in range(top, top + h):
                        for j in range(left, left + w):
                            if mask[i, j] == 0:
                                mask[i, j] = 1
                                delta += 1
                if delta > 0:
                    break
        return delta

    def __call__(self):  # :77-105
        mask = np.zeros(shape=(self.height, self.width), dtype=np.int32)
        mask_count = 0
        while mask_count < self.num_masking_patches:
            max_mask_patches = min(self.num_masking_patches - mask_count, self.max_num_patches)
            delta = self._mask(mask, max_mask_patches)
            if delta == 0:
                break
            mask_count += delta
        if mask_count > self.num_masking_patches:
            delta = mask_count - self.num_masking_patches
            mx, my = mask.nonzero()
            to_vis = np.random.choice(mx.shape[0], delta, replace=False)
            mask[mx[to_vis], my[to_vis]] = 0
        elif mask_count < self.num_masking_patches:
            delta = self.num_masking_patches - mask_count
            mx, my = (mask == 0).nonzero()
            to_mask = np.random.choice(mx.shape[0], delta, replace=False)
            mask[mx[to_mask], my[to_mask]] = 1
        assert mask.sum() == self.num_masking_patches
        return mask


def sample_mim_masks(cfg, B):
    """models/beit2.py:432-439 — B generator calls, stacked to bool [B, num_patches]."""
    gen = MaskingGenerator(cfg["image_res"] // cfg["patch_size"], cfg["num_masking_patches"],
                           cfg["min_num_patches"])
    rows = [torch.Tensor(gen().flatten()) for _ in range(B)]
    return torch.stack(rows, 0).bool()


# ----------------------------------------------------------------------------------------------
# vision encoder (BEiT-v2)
# ----------------------------------------------------------------------------------------------


def _ln(x, sd, name, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def beit_attention(x, sd, p, heads, rel_bias=None):
    """models/beit2.py:126-166 (also models/vqkd_vit.py:122-160 when rel_bias is None)."""
    B, N, Cdim = x.shape
    qkv_bias = torch.cat((sd[p + "q_bias"], torch.zeros_like(sd[p + "v_bias"]), sd[p + "v_bias"]))
    qkv = F.linear(x, sd[p + "qkv.weight"], qkv_bias)
    qkv = qkv.reshape(B, N, 3, heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (q.shape[-1] ** -0.5)
    attn = q @ k.transpose(-2, -1)
    if rel_bias is not None:
        attn = attn + rel_bias.unsqueeze(0)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def beit_rel_bias(sd, p, ws):
    """models/beit2.py:139-144 — gather the [H, N, N] bias from the table."""
    idx = relative_position_index(ws)
    n = ws * ws + 1
    return sd[p + "relative_position_bias_table"][idx.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()


def beit_block(x, sd, p, heads, eps, ws=None, layerscale=True):
    """models/beit2.py:191-206 (eval: DropPath is identity)."""
    rel = beit_rel_bias(sd, p + "attn.", ws) if ws is not None else None
    y = beit_attention(_ln(x, sd, p + "norm1", eps), sd, p + "attn.", heads, rel)
    x = x + (sd[p + "gamma_1"] * y if layerscale else y)
    h = F.linear(_ln(x, sd, p + "norm2", eps), sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = F.linear(F.gelu(h), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + (sd[p + "gamma_2"] * h if layerscale else h)


def patch_embed(image, sd, p, P):
    """models/beit2.py:223-230 — 16x16 stride-16 conv, flatten, transpose."""
    x = F.conv2d(image, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=P)
    return x.flatten(2).transpose(1, 2)


def vision_forward(image, sd, cfg, ids_mask=None, prefix="vision_encoder.", collect=None):
    """models/beit2.py:423-466 forward_avgpool: returns [B, 1+num_patches, D] (token 0 = mean of LN'd patches)."""
    x = patch_embed(image, sd, prefix, cfg["patch_size"])
    B, S, D = x.shape
    if ids_mask is not None:  # :438-443
        w = ids_mask.unsqueeze(-1).type_as(x)
        x = x * (1 - w) + sd[prefix + "mask_token"].expand(B, S, -1) * w
    x = torch.cat((sd[prefix + "cls_token"].expand(B, -1, -1), x), dim=1)
    ws = cfg["image_res"] // cfg["patch_size"]
    for i in range(cfg["vision_depth"]):
        x = beit_block(x, sd, f"{prefix}blocks.{i}.", cfg["vision_heads"], cfg["vision_ln_eps"], ws=ws)
        if collect is not None:
            collect.append(x)
    x = x[:, 1:]
    x = _ln(x, sd, prefix + "fc_norm", cfg["vision_ln_eps"])
    x_cls = x.mean(dim=1, keepdim=True)
    return torch.cat([x_cls, x], dim=1)


# ----------------------------------------------------------------------------------------------
# text / fusion encoder (RoBERTa, post-LN)
# ----------------------------------------------------------------------------------------------


def roberta_position_ids(input_ids, pad_id):
    """models/xroberta.py:1747-1757."""
    mask = input_ids.ne(pad_id).int()
    return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + pad_id


def bert_embeddings(input_ids, sd, p, cfg):
    """models/xbert.py:167-221 (eval): word + token_type[0] + position[0..L-1], LayerNorm."""
    e = p + "bert.embeddings."
    L = input_ids.shape[1]
    emb = F.embedding(input_ids, sd[e + "word_embeddings.weight"], padding_idx=cfg["pad_id"])   # xbert.py:172
    emb = emb + sd[e + "token_type_embeddings.weight"][0] + sd[e + "position_embeddings.weight"][:L]
    return _ln(emb, sd, e + "LayerNorm", cfg["ln_eps"])


def roberta_embeddings(input_ids, sd, p, cfg):
    """models/xroberta.py:104-137 (eval: dropout off)."""
    e = p + "roberta.embeddings."
    pos = roberta_position_ids(input_ids, cfg["pad_id"])
    emb = F.embedding(input_ids, sd[e + "word_embeddings.weight"], padding_idx=cfg["pad_id"])   # xroberta.py:80,100-102
    emb = emb + sd[e + "token_type_embeddings.weight"][0]
    emb = emb + F.embedding(pos, sd[e + "position_embeddings.weight"], padding_idx=cfg["pad_id"])
    return _ln(emb, sd, e + "LayerNorm", cfg["ln_eps"])


def _heads(x, H):
    B, L, D = x.shape
    return x.view(B, L, H, D // H).permute(0, 2, 1, 3)


def roberta_attention(h, ext_mask, sd, p, H, eps, enc=None, enc_mask=None, scale_after=False):
    """models/xroberta.py:201-289 (self / cross) + RobertaSelfOutput :300-304.  scale_after: models/xbert.py:296-301,329-330
    without config.fp16 divides the SCORES by sqrt(d) instead of q."""
    q = F.linear(h, sd[p + "self.query.weight"], sd[p + "self.query.bias"])
    src, mask = (h, ext_mask) if enc is None else (enc, enc_mask)
    k = _heads(F.linear(src, sd[p + "self.key.weight"], sd[p + "self.key.bias"]), H)
    v = _heads(F.linear(src, sd[p + "self.value.weight"], sd[p + "self.value.bias"]), H)
    d = q.shape[-1] // H
    q = _heads(q, H)
    if not scale_after:
        q = q / math.sqrt(d)  # scaled BEFORE QK^T (:237)
    s = torch.matmul(q, k.transpose(-1, -2))
    if scale_after:
        s = s / math.sqrt(d)
    if mask is not None:
        s = s + mask
    pr = torch.softmax(s, dim=-1)
    ctx = torch.matmul(pr, v).permute(0, 2, 1, 3).contiguous()
    ctx = ctx.view(ctx.shape[0], ctx.shape[1], -1)
    o = F.linear(ctx, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return _ln(o + h, sd, p + "output.LayerNorm", eps)


def roberta_layer(h, ext_mask, sd, p, cfg, enc=None, enc_mask=None, scale_after=False):
    """models/xroberta.py:405-473 (= models/xbert.py:455-533): self-attn -> [cross-attn] -> FFN, all post-LN."""
    a = roberta_attention(h, ext_mask, sd, p + "attention.", cfg["heads"], cfg["ln_eps"], scale_after=scale_after)
    if enc is not None:
        a = roberta_attention(a, ext_mask, sd, p + "crossattention.", cfg["heads"], cfg["ln_eps"], enc, enc_mask)
    i = F.gelu(F.linear(a, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    o = F.linear(i, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return _ln(o + a, sd, p + "output.LayerNorm", cfg["ln_eps"])


def extended_mask(atts):
    """models/xroberta.py:805-806 — additive (1-m)*-10000, broadcast to [B,1,1,L]."""
    return (1.0 - atts[:, None, None, :].to(torch.float32)) * -10000.0


def inverted_mask(atts):
    """HF invert_attention_mask (xroberta.py:903-909): (1-m)*finfo.min; exactly 0 for all-ones masks."""
    return (1.0 - atts[:, None, None, :].to(torch.float32)) * torch.finfo(torch.float32).min


def text_forward(text_ids, text_atts, sd, cfg, prefix="text_encoder.", collect=None):
    """XFMBase.get_text_embeds (models/xfm.py:600-611): 12 layers without cross-attention."""
    bert = cfg.get("text_arch", "roberta") == "bert" and prefix == "text_encoder."
    h = bert_embeddings(text_ids, sd, prefix, cfg) if bert else roberta_embeddings(text_ids, sd, prefix, cfg)
    m = extended_mask(text_atts)
    stem = "bert." if bert else "roberta."
    for i in range(cfg["text_layers"]):
        h = roberta_layer(h, m, sd, f"{prefix}{stem}encoder.layer.{i}.", cfg,
                          scale_after=bert and not cfg.get("text_fp16", True))
        if collect is not None:
            collect.append(h)
    return h


def fusion_forward(text_embeds, text_atts, image_embeds, image_atts, sd, cfg, prefix="fusion_encoder.", collect=None):
    """XFMBase.get_cross_embeds (models/xfm.py:659-680) with encoder_embeds (bypasses embeddings,
    xroberta.py:920-929); cross-attention in every layer."""
    h = text_embeds
    m = extended_mask(text_atts)
    em = inverted_mask(image_atts)
    for i in range(cfg["fusion_layers"]):
        h = roberta_layer(h, m, sd, f"{prefix}roberta.encoder.layer.{i}.", cfg, image_embeds, em)
        if collect is not None:
            collect.append(h)
    return h


def lm_head(x, sd, p, cfg):
    """RobertaLMHead (models/xroberta.py:1313-1333) / BertOnlyMLMHead (models/xbert.py:663-707), decoder tied to the word
    embeddings."""
    if cfg.get("text_arch", "roberta") == "bert" and p == "text_encoder.":
        c = p + "cls.predictions."
        x = F.linear(x, sd[c + "transform.dense.weight"], sd[c + "transform.dense.bias"])
        x = _ln(F.gelu(x), sd, c + "transform.LayerNorm", cfg["ln_eps"])
        return F.linear(x, sd[p + "bert.embeddings.word_embeddings.weight"], sd[c + "bias"])
    x = F.linear(x, sd[p + "lm_head.dense.weight"], sd[p + "lm_head.dense.bias"])
    x = _ln(F.gelu(x), sd, p + "lm_head.layer_norm", cfg["ln_eps"])
    return F.linear(x, sd[p + "roberta.embeddings.word_embeddings.weight"], sd[p + "lm_head.bias"])


def mlm_loss_from_hidden(seq, masked_pos, masked_ids, sd, p, cfg):
    """models/xroberta.py:1275-1299: gather masked positions, LM head, CE(ignore_index=-100)."""
    g = torch.gather(seq, 1, masked_pos.unsqueeze(2).expand(-1, -1, seq.size(-1)))
    logits = lm_head(g, sd, p, cfg)
    return F.cross_entropy(logits.view(-1, cfg["vocab_size"]), masked_ids.view(-1))


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------


def get_features(image_embeds, text_embeds, sd):
    """models/xfm.py:614-621."""
    i = F.normalize(F.linear(image_embeds[:, 0, :], sd["vision_proj.weight"], sd["vision_proj.bias"]), dim=-1)
    t = F.normalize(F.linear(text_embeds[:, 0, :], sd["text_proj.weight"], sd["text_proj.bias"]), dim=-1)
    return i, t


def contrastive_loss(image_feat_all, text_feat_all, temp, idx_all=None):
    """models/xfm.py:683-715 on already-gathered features ([B*W, E])."""
    logits = image_feat_all @ text_feat_all.t() / temp
    n = logits.shape[0]
    if idx_all is None:
        labels = torch.arange(n, device=logits.device)
        li = F.cross_entropy(logits, labels)
        lt = F.cross_entropy(logits.t(), labels)
    else:
        idx_all = idx_all.view(-1, 1)
        pos = torch.eq(idx_all, idx_all.t()).float()
        labels = pos / pos.sum(1, keepdim=True)
        li = -torch.sum(F.log_softmax(logits, dim=1) * labels, dim=1).mean()
        lt = -torch.sum(F.log_softmax(logits.t(), dim=1) * labels, dim=1).mean()
    return (li + lt) / 2


def hard_negative_weights(image_feat, text_feat, temp, idx=None):
    """models/xfm.py:717-734 — sampling weights (deterministic part of get_hard_negatives)."""
    sim_i2t = image_feat @ text_feat.t() / temp
    sim_t2i = text_feat @ image_feat.t() / temp
    w_i2t = F.softmax(sim_i2t, dim=1) + 1e-5
    w_t2i = F.softmax(sim_t2i, dim=1) + 1e-5
    if idx is None:
        w_i2t.fill_diagonal_(0)
        w_t2i.fill_diagonal_(0)
    else:
        idx = idx.view(-1, 1)
        mask = torch.eq(idx, idx.t())
        w_i2t.masked_fill_(mask, 0)
        w_t2i.masked_fill_(mask, 0)
    return w_i2t, w_t2i


def itm_head(x, sd, name="itm_head"):
    """build_mlp (models/xfm.py:115-121): Linear -> LayerNorm(1e-5) -> GELU -> Linear."""
    x = F.linear(x, sd[name + ".0.weight"], sd[name + ".0.bias"])
    x = F.gelu(F.layer_norm(x, (x.shape[-1],), sd[name + ".1.weight"], sd[name + ".1.bias"], 1e-5))
    return F.linear(x, sd[name + ".3.weight"], sd[name + ".3.bias"])


def matching_loss(image_embeds, image_atts, text_embeds, text_atts, image_neg_idx, text_neg_idx, sd, cfg,
                  is_pretrain=True, collect=None):
    """models/xfm.py:749-802 with the sampled negative indices supplied by the caller."""
    bs = image_embeds.shape[0]
    ine = image_embeds[image_neg_idx]
    ina = image_atts[image_neg_idx]
    tne = text_embeds[text_neg_idx]
    tna = text_atts[text_neg_idx]
    te_all = torch.cat([text_embeds, tne], 0)
    ta_all = torch.cat([text_atts, tna], 0)
    ie_all = torch.cat([ine, image_embeds], 0)
    ia_all = torch.cat([ina, image_atts], 0)
    det = (lambda t: t.detach()) if is_pretrain else (lambda t: t)
    cross_pos = fusion_forward(det(text_embeds), text_atts, image_embeds, image_atts, sd, cfg, collect=collect)[:, 0]
    cross_neg = fusion_forward(det(te_all), ta_all, ie_all, ia_all, sd, cfg)[:, 0]
    out = itm_head(torch.cat([cross_pos, cross_neg], 0), sd)
    labels = torch.cat([torch.ones(bs, dtype=torch.long), torch.zeros(2 * bs, dtype=torch.long)]).to(out.device)
    return F.cross_entropy(out, labels), cross_pos


def fuse_mlm_loss(text_ids_masked, text_atts, image_embeds, image_atts, masked_pos, masked_ids, sd, cfg):
    """models/xfm.py:638-656 (text embeds of the masked ids are detached)."""
    te = text_forward(text_ids_masked, text_atts, sd, cfg).detach()
    seq = fusion_forward(te, text_atts, image_embeds, image_atts, sd, cfg)
    return mlm_loss_from_hidden(seq, masked_pos, masked_ids, sd, "fusion_encoder.", cfg)


def text_mlm_loss(text_ids_masked, text_atts, masked_pos, masked_ids, sd, cfg):
    """models/xfm.py:805-812 on the text-only stream (image args None)."""
    seq = text_forward(text_ids_masked, text_atts, sd, cfg)
    return mlm_loss_from_hidden(seq, masked_pos, masked_ids, sd, "text_encoder.", cfg)


def mim_loss_mse(image_embeds_masked, targets, mask):
    """models/xfm.py:631-635 (mim_cls_only False)."""
    targets = targets.detach()
    a = F.mse_loss(image_embeds_masked[:, 1:, :][mask], targets[:, 1:, :][mask])
    return a + F.mse_loss(image_embeds_masked[:, 0, :], targets[:, 0, :])


# ----------------------------------------------------------------------------------------------
# VQ-KD tokenizer
# ----------------------------------------------------------------------------------------------


def quantizer_distances(z_flat, codebook):
    """models/norm_ema_quantizer.py:158-160."""
    return z_flat.pow(2).sum(dim=1, keepdim=True) + codebook.pow(2).sum(dim=1) - 2 * torch.einsum(
        "bd,nd->bn", z_flat, codebook)


def quantizer_indices(z, codebook):
    """models/norm_ema_quantizer.py:149-162: z [B, C, h, w] -> l2norm over C -> argmin (first index on ties)."""
    zf = F.normalize(z.permute(0, 2, 3, 1), p=2, dim=-1).reshape(-1, codebook.shape[1])
    return torch.argmin(quantizer_distances(zf, codebook), dim=1)


def quantizer_ambiguous(z, codebook, tol=1e-6):
    """Rows whose best and second-best distance differ by < tol in fp64: an fp32 implementation with a
    different summation order may legitimately pick either; parity tests exclude exactly these rows."""
    zf = F.normalize(z.permute(0, 2, 3, 1).double(), p=2, dim=-1).reshape(-1, codebook.shape[1])
    d = quantizer_distances(zf, codebook.double())
    top2 = torch.topk(d, 2, dim=1, largest=False).values
    return (top2[:, 1] - top2[:, 0]) < tol


def vqkd_preprocess(data):
    """models/model_vqkd.py:125-131 (process_type 'default')."""
    if data.max() <= 1.0:
        data = data * 255.0
    return data / 127.5 - 1.0


def vqkd_features(image, sd, cfg, prefix="vqkd."):
    """models/model_vqkd.py:151-160 up to the quantizer input: [B, code_dim, h, w]."""
    e = prefix + "encoder."
    x = patch_embed(image, sd, e, cfg["patch_size"])
    B = x.shape[0]
    x = torch.cat((sd[e + "cls_token"].expand(B, -1, -1), x), dim=1) + sd[e + "pos_embed"]
    D = x.shape[-1]
    for i in range(cfg["vision_depth"]):
        x = beit_block(x, sd, f"{e}blocks.{i}.", D // 64, 1e-6, ws=None, layerscale=False)
    x = _ln(x[:, 1:, :], sd, e + "fc_norm", 1e-6)  # vqkd_vit.py:391-397 return_patch_tokens
    t = prefix + "encode_task_layer."
    x = F.linear(torch.tanh(F.linear(x, sd[t + "0.weight"], sd[t + "0.bias"])), sd[t + "2.weight"], sd[t + "2.bias"])
    h = int(math.sqrt(x.shape[1]))
    return x.transpose(1, 2).reshape(B, -1, h, h)


def vqkd_codebook_indices(image, sd, cfg, prefix="vqkd."):
    """models/model_vqkd.py:173-175 get_codebook_indices -> [B, num_patches] int64."""
    z = vqkd_features(vqkd_preprocess(image), sd, cfg, prefix)
    return quantizer_indices(z, sd[prefix + "quantize.embedding.weight"]).view(image.shape[0], -1)


def mim_loss_vqkd(image_embeds_masked, image, mask, sd, cfg):
    """models/xfm.py:625-629."""
    with torch.no_grad():
        ids = vqkd_codebook_indices(image, sd, cfg)
    logits = F.linear(image_embeds_masked[:, 1:, :][mask], sd["lm_head.weight"], sd["lm_head.bias"])
    return F.cross_entropy(logits, ids[mask])


# ----------------------------------------------------------------------------------------------
# VQA: cross-modal causal decoder (models/model_generation.py:23-202, models/xroberta.py:963-1123)
# ----------------------------------------------------------------------------------------------


def causal_extended_mask(atts):
    """models/xroberta.py:771-806 with is_decoder=True: (1 - causal[i,j] * atts[b,j]) * -10000, [B,1,L,L]."""
    L = atts.shape[1]
    ids = torch.arange(L)
    causal = (ids[None, :] <= ids[:, None]).to(torch.float32)  # [i, j]: key j visible from query i
    ext = causal[None, None, :, :] * atts[:, None, None, :].to(torch.float32)
    return (1.0 - ext) * -10000.0


def decoder_forward(input_ids, atts, enc_states, enc_atts, sd, cfg, prefix="text_decoder."):
    """RobertaForCausalLM.forward up to the LM head (models/xroberta.py:1079-1097): embeddings, then every layer =
    causal self-attention -> cross-attention over `enc_states` (mask: invert_attention_mask, :903-909) -> FFN.
    Returns the vocabulary logits [B, L, V]."""
    h = roberta_embeddings(input_ids, sd, prefix, cfg)
    m = causal_extended_mask(atts)
    em = inverted_mask(enc_atts)
    for i in range(cfg["dec_layers"]):
        h = roberta_layer(h, m, sd, f"{prefix}roberta.encoder.layer.{i}.", cfg, enc_states, em)
    return lm_head(h, sd, prefix, cfg)


def causal_lm_loss(logits, labels):
    """models/xroberta.py:1104-1110 with reduction='none': next-token CE (ignore_index -100), summed per sequence."""
    V = logits.shape[-1]
    loss = F.cross_entropy(logits[:, :-1, :].reshape(-1, V), labels[:, 1:].reshape(-1), reduction="none")
    return loss.view(logits.shape[0], -1).sum(1)


def vqa_question_states(image, q_ids, q_atts, sd, cfg):
    """models/model_generation.py:94-109: vision encoder, text encoder, fusion encoder (is_pretrain=False)."""
    image_embeds = vision_forward(image, sd, cfg)
    image_atts = torch.ones(image_embeds.shape[:-1], dtype=torch.long)
    text_embeds = text_forward(q_ids, q_atts, sd, cfg)
    return fusion_forward(text_embeds, q_atts, image_embeds, image_atts, sd, cfg)


def vqa_train_loss(image, q_ids, q_atts, a_ids, a_atts, k, weights, sd, cfg, collect=None):
    """XFMForVQA.forward(train=True) (models/model_generation.py:96-131): every answer of question b attends to
    question_output[b]; loss = sum(weights * per-answer loss) / batch."""
    q_out = vqa_question_states(image, q_ids, q_atts, sd, cfg)
    rep = torch.repeat_interleave(torch.arange(len(k)), torch.as_tensor(k))
    targets = a_ids.masked_fill(a_ids == cfg["pad_id"], -100)
    logits = decoder_forward(a_ids, a_atts, q_out[rep], q_atts[rep], sd, cfg)
    per_answer = causal_lm_loss(logits, targets)
    if collect is not None:
        collect.update(question_output=q_out, answer_loss=per_answer, logits=logits)
    return (weights * per_answer).sum() / image.shape[0]


def vqa_rank_answer(q_states, q_atts, answer_ids, answer_atts, k, sd, cfg):
    """XFMForVQA.rank_answer (models/model_generation.py:146-202): first-token probabilities of every candidate,
    top-k, full-sequence log-likelihood of those k, softmax re-rank."""
    nq = q_states.shape[0]
    start = answer_ids[0, 0].repeat(nq, 1)
    logits = decoder_forward(start, torch.ones_like(start), q_states, q_atts, sd, cfg)[:, 0, :]
    p_first = torch.softmax(logits, dim=1).index_select(1, answer_ids[:, 1])
    topk_probs, topk_ids = p_first.topk(k, dim=1)
    ids = answer_ids[topk_ids.reshape(-1)]
    atts = answer_atts[topk_ids.reshape(-1)]
    targets = ids.masked_fill(ids == cfg["pad_id"], -100)
    rep = torch.repeat_interleave(torch.arange(nq), k)
    loss = causal_lm_loss(decoder_forward(ids, atts, q_states[rep], q_atts[rep], sd, cfg), targets)
    log_probs = torch.cat([topk_probs.view(-1, 1).log(), -loss.view(-1, 1)], dim=1).sum(1).view(nq, k)
    probs, rerank = torch.softmax(log_probs, dim=-1).topk(k, dim=1)
    return torch.gather(topk_ids, 1, rerank), probs


def vqa_rank(image, q_ids, q_atts, answer_ids, answer_atts, k, sd, cfg):
    """XFMForVQA.forward(train=False) (models/model_generation.py:133-144); question_atts is all-ones there (:141)."""
    q_out = vqa_question_states(image, q_ids, q_atts, sd, cfg)
    return vqa_rank_answer(q_out, torch.ones(q_out.shape[:-1], dtype=torch.long), answer_ids, answer_atts, k, sd, cfg)


def make_vqa_batch(cfg, B=3, L=12, La=6, n_cand=7, seed=11):
    """Synthetic VQA batch: questions (odd rows padded), 1-3 answers per question with weights, a candidate list."""
    g = torch.Generator().manual_seed(seed)
    base = make_batch(cfg, B, L=L, M=2, seed=seed)
    V = cfg["vocab_size"]
    k = [(b % 3) + 1 for b in range(B)]

    def answers(n):
        ids = torch.randint(3, V - 1, (n, La), generator=g)
        ids[:, 0] = 0
        atts = torch.ones(n, La, dtype=torch.long)
        for a in range(n):
            n_real = 3 + (a % (La - 2))        # 3 .. La tokens incl. BOS / EOS
            ids[a, n_real - 1] = 2
            ids[a, n_real:] = cfg["pad_id"]
            atts[a, n_real:] = 0
        return ids, atts
    a_ids, a_atts = answers(sum(k))
    weights = torch.rand(sum(k), generator=g) * 0.8 + 0.2
    c_ids, c_atts = answers(n_cand)
    return dict(image=base["image"], q_ids=base["text_ids"], q_atts=base["text_atts"], a_ids=a_ids, a_atts=a_atts, k=k,
                weights=weights, cand_ids=c_ids, cand_atts=c_atts)


# ----------------------------------------------------------------------------------------------
# region / bounding-box branch (models/xfm.py:574-597,815-854, models/beit2.py:468-475, models/box_ops.py)
# ----------------------------------------------------------------------------------------------


def vision_forward_region(image, idx_to_group_img, image_atts, sd, cfg):
    """beit2.py:468-475 + xfm.py:588-597: (region-pooled embeds [bsz, 1+np, D], full-image embeds gathered to [bsz, ...])."""
    full = vision_forward(image, sd, cfg)
    x = full[:, 1:]
    x_bs = x[idx_to_group_img]
    weights = image_atts[:, 1:].unsqueeze(2)
    x_bs_cls = torch.sum(weights * x_bs, dim=1, keepdim=True) / torch.sum(weights, dim=1, keepdim=True)
    return torch.cat([x_bs_cls, x_bs], dim=1), full[idx_to_group_img]


def box_cxcywh_to_xyxy(x):
    """box_ops.py:10-14."""
    x_c, y_c, w, h = x.unbind(-1)
    return torch.stack([x_c - 0.5 * w, y_c - 0.5 * h, x_c + 0.5 * w, y_c + 0.5 * h], dim=-1)


def generalized_box_iou(boxes1, boxes2):
    """box_ops.py:25-58 (pairwise [N, M])."""
    area1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    area2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    lt = torch.max(boxes1[:, None, :2], boxes2[:, :2])
    rb = torch.min(boxes1[:, None, 2:], boxes2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    union = area1[:, None] + area2 - inter
    iou = inter / union
    lt = torch.min(boxes1[:, None, :2], boxes2[:, :2])
    rb = torch.max(boxes1[:, None, 2:], boxes2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    area = wh[:, :, 0] * wh[:, :, 1]
    return iou - (area - union) / area


def bbox_loss(output_coord, target_bbox, is_image=None):
    """models/xfm.py:815-840."""
    loss_bbox = F.l1_loss(output_coord, target_bbox, reduction="none")
    boxes1, boxes2 = box_cxcywh_to_xyxy(output_coord), box_cxcywh_to_xyxy(target_bbox)
    if (boxes1[:, 2:] < boxes1[:, :2]).any() or (boxes2[:, 2:] < boxes2[:, :2]).any():
        loss_giou = torch.zeros(output_coord.size(0), device=output_coord.device)
    else:
        loss_giou = 1 - torch.diag(generalized_box_iou(boxes1, boxes2))
    if is_image is None:
        num_boxes = target_bbox.size(0)
    else:
        num_boxes = torch.sum(1 - is_image)
        loss_bbox = loss_bbox * (1 - is_image.view(-1, 1))
        loss_giou = loss_giou * (1 - is_image)
    return loss_bbox.sum() / num_boxes, loss_giou.sum() / num_boxes


def predict_bbox(image_embeds_full, text_embeds, text_atts, sd, cfg, is_pretrain=True):
    """models/xfm.py:843-854."""
    ones = torch.ones(image_embeds_full.shape[:2], dtype=torch.long, device=image_embeds_full.device)
    te = text_embeds.detach() if is_pretrain else text_embeds
    cls = fusion_forward(te, text_atts, image_embeds_full, ones, sd, cfg)[:, 0, :]
    return itm_head(cls, sd, name="bbox_head").sigmoid()


def pretrain_forward_region(sd, cfg, batch, image_neg_idx, text_neg_idx, collect=None):
    """models/model_pretrain.py:30-91 on a region batch (ret_bbox_loss = ret_bbox_giou = True, no MIM): batch carries
    idx_to_group_img [bsz], image_atts [bsz, 1+np], target_bbox [bsz, 4], is_image [bsz] besides the text tensors."""
    temp = sd["temp"].clamp(cfg["min_temp"], cfg["max_temp"])
    c = collect if collect is not None else {}
    image_atts = batch["image_atts"]
    image_embeds, image_full = vision_forward_region(batch["image"], batch["idx_to_group_img"], image_atts, sd, cfg)
    text_embeds = text_forward(batch["text_ids"], batch["text_atts"], sd, cfg)
    image_feat, text_feat = get_features(image_embeds, text_embeds, sd)
    c["image_embeds"], c["image_embeds_fullatts"], c["image_feat"] = image_embeds, image_full, image_feat
    loss_itc = contrastive_loss(image_feat, text_feat, temp)
    loss_itm, _ = matching_loss(image_embeds, image_atts, text_embeds, batch["text_atts"], image_neg_idx, text_neg_idx, sd, cfg)
    loss_mlm = fuse_mlm_loss(batch["text_ids_masked"], batch["text_atts"], image_embeds, image_atts, batch["masked_pos"],
                             batch["masked_ids"], sd, cfg)
    coord = predict_bbox(image_full, text_embeds, batch["text_atts"], sd, cfg)
    c["output_coord"] = coord
    loss_bbox, loss_giou = bbox_loss(coord, batch["target_bbox"], batch.get("is_image"))
    return dict(loss_itc=loss_itc, loss_itm=loss_itm, loss_mlm=loss_mlm, loss_bbox=loss_bbox, loss_giou=loss_giou)


def make_region_batch(cfg, n_img=3, bsz=6, L=24, M=6, seed=31):
    """Synthetic region batch in the layout of dataset/pretrain_dataset.py's region collate (Pretrain.py:93-99): several
    samples per image, rectangular region masks over the patch grid (position 0 = cls, always visible), boxes in cxcywh."""
    base = make_batch(cfg, bsz, L=L, M=M, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    ws = cfg["image_res"] // cfg["patch_size"]
    image = torch.randn(n_img, 3, cfg["image_res"], cfg["image_res"], generator=g)
    idx = torch.arange(bsz) * n_img // bsz
    atts = torch.zeros(bsz, ws * ws + 1, dtype=torch.long)
    atts[:, 0] = 1
    target = torch.zeros(bsz, 4)
    is_image = torch.zeros(bsz, dtype=torch.long)
    for b in range(bsz):
        if b % 3 == 2:   # a whole-image caption among the region samples
            atts[b] = 1
            target[b] = torch.tensor([0.5, 0.5, 1.0, 1.0])
            is_image[b] = 1
            continue
        x0, y0 = int(torch.randint(0, ws - 1, (1,), generator=g)), int(torch.randint(0, ws - 1, (1,), generator=g))
        x1, y1 = int(torch.randint(x0 + 1, ws + 1, (1,), generator=g)), int(torch.randint(y0 + 1, ws + 1, (1,), generator=g))
        grid = torch.zeros(ws, ws, dtype=torch.long)
        grid[y0:y1, x0:x1] = 1
        atts[b, 1:] = grid.reshape(-1)
        target[b] = torch.tensor([(x0 + x1) / 2 / ws, (y0 + y1) / 2 / ws, (x1 - x0) / ws, (y1 - y0) / ws])
    out = dict(base)
    out.update(image=image, idx_to_group_img=idx, image_atts=atts, target_bbox=target, is_image=is_image)
    return out


# ----------------------------------------------------------------------------------------------
# whole pre-training forward (models/model_pretrain.py:30-91)
# ----------------------------------------------------------------------------------------------


def pretrain_forward(sd, cfg, batch, image_neg_idx, text_neg_idx, ids_mask=None, collect=None, world_feats=None):
    """ITC + ITM + MLM (+ MIM) in eval mode.  `collect`, if a dict, receives per-layer activations.
    `world_feats`: optional (image_feat_others, text_feat_others) from other ranks appended AFTER the local
    slice in rank order is the caller's business — single-process callers leave it None."""
    temp = sd["temp"].clamp(cfg["min_temp"], cfg["max_temp"])
    c = collect if collect is not None else {}
    c["vision"], c["text"], c["fusion_pos"], c["vision_masked"] = [], [], [], []
    image = batch["image"]
    image_embeds = vision_forward(image, sd, cfg, collect=c["vision"])
    image_atts = torch.ones(image_embeds.shape[:-1], dtype=torch.long, device=image_embeds.device)
    text_embeds = text_forward(batch["text_ids"], batch["text_atts"], sd, cfg, collect=c["text"])
    image_feat, text_feat = get_features(image_embeds, text_embeds, sd)
    c["image_embeds"], c["text_embeds"], c["image_feat"], c["text_feat"] = image_embeds, text_embeds, image_feat, text_feat
    loss_itc = contrastive_loss(image_feat, text_feat, temp)
    w_i2t, w_t2i = hard_negative_weights(image_feat.detach(), text_feat.detach(), temp.detach())
    c["weights_i2t"], c["weights_t2i"] = w_i2t, w_t2i
    loss_itm, cross_pos = matching_loss(image_embeds, image_atts, text_embeds, batch["text_atts"], image_neg_idx,
                                        text_neg_idx, sd, cfg, collect=c["fusion_pos"])
    c["cross_pos"] = cross_pos
    loss_mlm = fuse_mlm_loss(batch["text_ids_masked"], batch["text_atts"], image_embeds, image_atts,
                             batch["masked_pos"], batch["masked_ids"], sd, cfg)
    out = dict(loss_itc=loss_itc, loss_itm=loss_itm, loss_mlm=loss_mlm)
    if ids_mask is not None:
        iem = vision_forward(image, sd, cfg, ids_mask=ids_mask, collect=c["vision_masked"])
        c["image_embeds_masked"] = iem
        if cfg["use_vision_tokenizer"]:
            out["loss_mim"] = mim_loss_vqkd(iem, image, ids_mask, sd, cfg)
        else:
            out["loss_mim"] = mim_loss_mse(iem, image_embeds, ids_mask)
    return out
