"""Configuration of the B200 path: the reference's flat YAML dict (configs/xfm-pt/Pretrain_XBrain_base_4m.yaml,
configs/model/config_beit2_base.json, <text_encoder>/config.json) normalised to one internal dict.

The reference reads these keys in models/xfm.py:124-130,205-256 (vision), :258-300 (text) and :471-539
(XFMBase.__init__).  Missing files (the GPU box has no ../data/roberta-base) fall back to the roberta-base /
BEiT-v2-base dimensions the shipped configs select.
"""
import json
import os

_DEFAULTS = dict(
    image_res=224, patch_size=16, vision_width=768, vision_depth=12, vision_heads=12, vision_mlp=3072,
    vision_ln_eps=1e-6, init_values=0.1, drop_path_rate=0.1,
    vocab_size=50265, hidden=768, text_layers=12, fusion_layers=12, heads=12, ffn=3072, max_pos=514,
    type_vocab=1, pad_id=1, ln_eps=1e-5, hidden_dropout=0.1, attn_dropout=0.1,
    embed_dim=256, temp=0.07, learnable_temp=True, min_temp=0.001, max_temp=0.5,
    num_masking_patches=75, min_num_patches=16,
    use_vision_tokenizer=False, codebook_size=8192, codebook_dim=32,
    use_bbox=True, detach_text_forMLM=True, mim_cls_only=False,
    text_arch="roberta",   # "bert": models/xbert.py text encoder (selected by a text_encoder path without "roberta", xfm.py:260-265)
)

# keys of the internal dict that a caller may pass directly (tests / bench use them for the tiny configuration)
_INTERNAL = set(_DEFAULTS)


def _read_json(path):
    if path and os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def normalize_config(config):
    cfg = dict(_DEFAULTS)
    config = dict(config or {})
    for k in _INTERNAL:
        if k in config:
            cfg[k] = config[k]
    vis = _read_json(config.get("vision_config"))
    if vis:
        cfg["vision_width"] = vis.get("vision_width", cfg["vision_width"])
        cfg["patch_size"] = vis.get("patch_size", cfg["patch_size"])
    if "text_encoder" in config:
        te = str(config["text_encoder"])
        if "roberta" in te:     # xfm.py:260-265
            cfg["text_arch"] = "roberta"
        elif "bert" in te:
            cfg["text_arch"] = "bert"
        else:
            raise ValueError(f"text_encoder {te!r}: neither a roberta nor a bert checkpoint directory")
        txt = _read_json(os.path.join(config["text_encoder"], "config.json"))
        if txt:
            cfg.update(vocab_size=txt["vocab_size"], hidden=txt["hidden_size"], heads=txt["num_attention_heads"],
                       ffn=txt["intermediate_size"], max_pos=txt["max_position_embeddings"],
                       type_vocab=txt["type_vocab_size"],
                       pad_id=txt.get("pad_token_id", 1 if cfg["text_arch"] == "roberta" else 0),
                       ln_eps=txt["layer_norm_eps"], hidden_dropout=txt["hidden_dropout_prob"],
                       attn_dropout=txt["attention_probs_dropout_prob"])
    if "text_num_hidden_layers" in config:
        cfg["text_layers"] = config["text_num_hidden_layers"]
    if "fusion_num_hidden_layers" in config:
        cfg["fusion_layers"] = config["fusion_num_hidden_layers"]
    if config.get("text_fusion_start_at", cfg["text_layers"]) != cfg["text_layers"]:
        raise NotImplementedError("xfm_b200 builds the shipped layout: text encoder without cross-attention "
                                  "(text_fusion_start_at == text_num_hidden_layers)")
    if config.get("fusion_fusion_start_at", 0) != 0:
        raise NotImplementedError("xfm_b200 builds the shipped layout: cross-attention in every fusion layer")
    if int(config.get("local_attn_depth", -1) or -1) > 0:
        raise NotImplementedError("local_attn_depth > 0 (beit2.py forward_localattn) is not built: every shipped yaml sets -1, "
                                  "region features come from the weighted average pool of beit2.py:468-475")
    if config.get("use_beit_v2", True) is not True:
        raise NotImplementedError("only the BEiT-v2 vision encoder (use_beit_v2: True) is built")
    if cfg["vision_width"] % 64 or cfg["hidden"] % 64 or cfg["vision_width"] // cfg["vision_heads"] != 64 or \
            cfg["hidden"] // cfg["heads"] != 64:
        raise NotImplementedError("attention kernels are built for head_dim 64")
    return cfg
