"""Checkpoint import for xfm_b200.XFMBase: the key-mapping and resolution-change rules of the reference's loaders,
producing state_dicts in the reference key layout that XFMBase.load_state_dict consumes.

    beit2_init_state        models/beit2.py:572-673   load_pretrained_beit2 (BEiT-v2 checkpoint -> vision_encoder.*)
    interpolate_rel_pos     models/beit2.py:611-652 / 753-808  relative-position table for another window size
    text_init_state         models/xfm.py:298-385      text encoder (RoBERTa or BERT class) from <text_encoder>/pytorch_model.bin
    finetune_state          models/xfm.py:408-468     load_pretrained(): XFM checkpoint -> fine-tuning model (BEiT-v2 branch)
    vqkd_state              models/model_vqkd.py:315-333       tokenizer weight file -> vqkd.*

Host-side, one-off work (numpy / scipy on the CPU); nothing here runs inside a training step.
"""
import os

import numpy as np
import torch


def interpolate_rel_pos(table, dst_size):
    """beit2.py:611-652: `table` [(2S-1)^2 + 3, H] for a window of side S' = 2S-1 relative offsets -> the table for
    dst_size relative offsets per axis.  Source offsets sit on a geometric progression (ratio q solved by bisection so that
    the progression spans dst_size // 2), the target ones on the integer grid; each head's (S' x S') surface is resampled
    with an interpolating bicubic spline.  The reference calls scipy.interpolate.interp2d(x, y, z, kind='cubic'), which on a
    rectilinear grid is FITPACK's regrid_smth with s = 0; RectBivariateSpline(y, x, z, kx=3, ky=3, s=0) is the same routine
    (scipy's interp2d migration guide) and is what the installed scipy still ships."""
    from scipy.interpolate import RectBivariateSpline

    num_extra = 3
    src_num, heads = table.shape
    src_size = int((src_num - num_extra) ** 0.5)
    if src_size == dst_size:
        return table
    extra = table[-num_extra:, :]
    body = table[:-num_extra, :]

    def geometric_progression(a, r, n):
        return a * (1.0 - r ** n) / (1.0 - r)

    left, right = 1.01, 1.5
    while right - left > 1e-6:
        q = (left + right) / 2.0
        gp = geometric_progression(1, q, src_size // 2)
        if gp > dst_size // 2:
            right = q
        else:
            left = q
    dis, cur = [], 1
    for i in range(src_size // 2):
        dis.append(cur)
        cur += q ** (i + 1)
    r_ids = [-d for d in reversed(dis)]
    x = np.asarray(r_ids + [0] + dis, dtype=np.float64)
    t = dst_size // 2.0
    dx = np.arange(-t, t + 0.1, 1.0)
    out = []
    for h in range(heads):
        z = body[:, h].reshape(src_size, src_size).float().numpy().astype(np.float64)
        f = RectBivariateSpline(x, x, z, kx=3, ky=3, s=0)   # z[j, i] = value at (y[j], x[i]); x == y here
        out.append(torch.tensor(f(dx, dx), dtype=torch.float32).contiguous().view(-1, 1))
    new = torch.cat(out, dim=-1).to(table.device)
    return torch.cat((new, extra.to(torch.float32)), dim=0)


def _interpolate_vision_tables(state, own_shapes, prefix=""):
    """beit2.py:753-808 (interpolate_pos_embed) / 604-652: drop relative_position_index buffers, resample every
    relative_position_bias_table whose size differs from the model's."""
    for key in list(state.keys()):
        if "relative_position_index" in key:
            state.pop(key)
            continue
        if "relative_position_bias_table" in key and (prefix + key) in own_shapes:
            dst_num = own_shapes[prefix + key][0]
            dst_size = int((dst_num - 3) ** 0.5)
            if state[key].shape[0] != dst_num:
                src_size = int((state[key].shape[0] - 3) ** 0.5)
                print("Position interpolate for %s from %dx%d to %dx%d" % (key, src_size, src_size, dst_size, dst_size))
                state[key] = interpolate_rel_pos(state[key], dst_size)
    return state


def beit2_init_state(ckpt_path, own_shapes, depth):
    """beit2.py:572-673: BEiT-v2 checkpoint file -> {'vision_encoder.<k>': tensor}.  own_shapes: name -> shape of this
    model's state_dict (to know the target table size)."""
    print("Load BEIT-V2 ckpt from %s" % ckpt_path)
    checkpoint = torch.load(ckpt_path, map_location="cpu")
    model = None
    for key in ("model", "module"):
        if key in checkpoint:
            model = checkpoint[key]
            print("Load state_dict by model_key = %s" % key)
            break
    if model is None:
        model = checkpoint
    model = dict(model)
    for k in ("head.weight", "head.bias"):
        model.pop(k, None)
    shared = "rel_pos_bias.relative_position_bias_table"
    if shared in model:
        print("Expand the shared relative position embedding to each transformer block. ")
        t = model.pop(shared)
        for i in range(depth):
            model["blocks.%d.attn.relative_position_bias_table" % i] = t.clone()
    model = _interpolate_vision_tables(model, own_shapes, prefix="vision_encoder.")
    return {"vision_encoder." + k: v for k, v in model.items()}


LAYER_PICK_24_TO_12 = {1: 0, 3: 1, 5: 2, 7: 3, 9: 4, 11: 5, 13: 6, 15: 7, 17: 8, 19: 9, 21: 10, 23: 11}  # xfm.py:309,313


def choose_layers(prefix, state, mapper):
    """xfm.py:64-78 load_params_choose_layers: keep layers listed in `mapper` (renumbered), drop the others."""
    for k in list(state.keys()):
        if k.startswith(prefix):
            new_k = None
            for i in mapper:
                if k.startswith(f"{prefix}.{i}."):
                    new_k = k.replace(f"{prefix}.{i}.", f"{prefix}.{mapper[i]}.")
                    break
            v = state.pop(k)
            if new_k:
                state[new_k] = v
    return state


def rename_tf_layernorm(state):
    """xfm.py:53-61: TensorFlow-era BERT checkpoints call the LayerNorm parameters gamma / beta."""
    for k in list(state.keys()):
        if "LayerNorm." in k:
            new_k = k.strip().replace("LayerNorm.beta", "LayerNorm.bias").replace("LayerNorm.gamma", "LayerNorm.weight")
            if new_k != k:
                state[new_k] = state.pop(k)


def text_init_state(text_encoder_dir, num_layers, arch="roberta"):
    """xfm.py:298-385: <dir>/pytorch_model.bin (a RobertaForMaskedLM checkpoint: roberta.*, lm_head.*; or a BertForMaskedLM one:
    bert.*, cls.predictions.*) -> the same keys under 'text_encoder.'.  Layer-picking rules of the layouts this package
    builds (text encoder without cross-attention): roberta-large / xlm-roberta-large and bert-large-uncased contribute every
    second layer when 12 layers are built; bert-base-uncased with 6 layers takes layers 1, 3, ... 11."""
    path = os.path.join(text_encoder_dir, "pytorch_model.bin")
    print("### Initializing text encoder from ", path)
    state = dict(torch.load(path, map_location="cpu"))
    if arch == "roberta":
        if ("roberta-large" in text_encoder_dir) and num_layers == 12:   # covers xlm-roberta-large too (xfm.py:303-314)
            choose_layers("roberta.encoder.layer", state, LAYER_PICK_24_TO_12)
    else:
        if "bert-large-uncased" in text_encoder_dir and "-12l" not in text_encoder_dir and "-18l" not in text_encoder_dir:
            rename_tf_layernorm(state)                                   # xfm.py:353-358
            if num_layers == 12:
                choose_layers("bert.encoder.layer", state, LAYER_PICK_24_TO_12)
        elif "bert-base-uncased" in text_encoder_dir and "-6l" not in text_encoder_dir and "-18l" not in text_encoder_dir:
            rename_tf_layernorm(state)                                   # xfm.py:326-330
            if num_layers == 6:
                choose_layers("bert.encoder.layer", state, {1: 0, 3: 1, 5: 2, 7: 3, 9: 4, 11: 5})
        elif "chinese-roberta-wwm-ext" in text_encoder_dir and num_layers == 6:   # xfm.py:371-374 (a BERT-class checkpoint)
            choose_layers("bert.encoder.layer", state, {1: 0, 3: 1, 5: 2, 7: 3, 9: 4, 11: 5})
    return {"text_encoder." + k: v for k, v in state.items()}


def finetune_state(state, own_shapes, text_encoder_name="roberta", load_text=True):
    """xfm.py:408-468 (BEiT-v2 branch): resample the vision tables to the model's resolution and, with load_text, strip
    the 'roberta.' segment from text_encoder keys (fine-tuning models hold a bare RobertaModel)."""
    state = dict(state)
    vis = {k[len("vision_encoder."):]: state.pop(k) for k in list(state.keys()) if k.startswith("vision_encoder.")}
    vis = _interpolate_vision_tables(vis, own_shapes, prefix="vision_encoder.")
    for k, v in vis.items():
        state["vision_encoder." + k] = v
    if load_text:
        name = "roberta." if "roberta" in text_encoder_name else "bert."
        for key in list(state.keys()):
            if key.startswith("text_encoder.") and name in key:
                state[key.replace(name, "")] = state.pop(key)
    return state


def vqkd_state(weight_path):
    """model_vqkd.py:315-333: tokenizer weight file -> {'vqkd.<k>'} for the parts the tokenizer path runs (encoder, task
    layer, codebook); decoder / teacher / loss entries are dropped."""
    w = torch.load(weight_path, map_location="cpu")
    w = w["model"] if "model" in w else w["state_dict"]
    keep = ("encoder.", "encode_task_layer.", "quantize.embedding.weight")
    return {"vqkd." + k: v for k, v in w.items() if k.startswith(keep)}
