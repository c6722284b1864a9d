"""XFMBase — the reference's module API (models/xfm.py:471-812) over the B200 kernel library.

Same constructor arguments, method names, argument meaning and state_dict keys as the reference class, so
models/model_pretrain.py / model_retrieval.py / model_nlvr.py-style subclasses and Pretrain.py-style drivers work
against it.  Underneath, nothing is an nn.Linear: all parameters are views into one flat fp32 buffer
(params.FlatParams), every method runs an explicit kernel schedule (encoders.py / blocks.py) and is exposed to
PyTorch autograd as ONE node per API call whose backward runs the hand-written backward schedule and accumulates
parameter gradients straight into the flat gradient buffer.

Differences a caller can observe (all documented in DESIGN.md):
  * tensors returned at the boundary are fp32 (like the reference) but computed in bf16 with fp32 accumulation;
  * get_hard_negatives returns two int64 device tensors (sampled on the GPU) instead of python lists — no host sync;
  * parameters that took no part in a backward pass keep grad None, like the reference's unused parameters;
  * get_matching_loss(..., return_cross_embeds=True) returns the positive pairs' CLS rows detached (no reference caller
    back-propagates through them);
  * get_text_embeds takes a private `_also_masked` argument and get_matching_and_fuse_mlm_loss exists besides the
    reference's two separate methods: schedule-level fusions used by xfm_b200/model_pretrain.py, results identical.
"""
import math
import os

import torch
import torch.nn as nn

from . import blocks as BK
from . import checkpoint as CK
from . import encoders as E
from . import lib as L
from .config import normalize_config
from .masking import BlockMaskSampler, sample_batch
from .params import FlatParams


def relative_position_index(ws):
    """Pair-wise relative position index of a ws x ws window plus the three cls entries (beit2.py:92-109)."""
    n_rel = (2 * ws - 1) ** 2 + 3
    ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    pos = torch.stack([ys.reshape(-1), xs.reshape(-1)], 1)  # [ws*ws, 2]
    d = pos[:, None, :] - pos[None, :, :] + (ws - 1)
    idx = torch.full((ws * ws + 1, ws * ws + 1), n_rel - 1, dtype=torch.int64)
    idx[1:, 1:] = d[..., 0] * (2 * ws - 1) + d[..., 1]
    idx[0, 1:] = n_rel - 3
    idx[1:, 0] = n_rel - 2
    return idx


def default_init(seed=0):
    """Random init in the spirit of the reference (trunc-normal .02 weights, zero biases, unit LayerNorm; beit2.py:340-353,
    xroberta.py _init_weights).  Tests and the bench overwrite it with their own synthetic weights."""
    g = torch.Generator().manual_seed(seed)

    def init(name, shape):
        leaf = name.rsplit(".", 1)[-1]
        if leaf in ("gamma_1", "gamma_2"):
            return torch.full(shape, 0.1)
        if ("norm" in name.lower() and leaf == "weight") or (name.endswith(".1.weight") and len(shape) == 1):
            return torch.ones(shape)
        if len(shape) <= 1 and leaf != "relative_position_bias_table":
            return torch.zeros(shape)
        if "quantize.embedding" in name:
            return torch.nn.functional.normalize(torch.randn(shape, generator=g), dim=-1)
        return torch.nn.init.trunc_normal_(torch.empty(shape), std=0.02, generator=g)
    return init


class _Holder(nn.Module):
    """Name-space node so that state_dict keys equal the reference's dotted names."""

    def forward(self, *a, **k):
        raise RuntimeError("xfm_b200 sub-modules are parameter holders; call the XFMBase methods")


class _HeadModule(_Holder):
    """Callable like the reference's nn.Sequential heads built by build_mlp (Retrieval.py:147-150 calls model.itm_head(x),
    model_nlvr.py:42 self.cls_head(x)); differentiable wrt x and the head's parameters."""

    def forward(self, x):
        return self._owner()._head_apply(self._head_name, x)


class _Op(torch.autograd.Function):
    """One autograd node per API call.  `impl` provides fwd(ctx, *tensors) -> tuple and bwd(ctx, *grads) -> tuple."""

    @staticmethod
    def forward(ctx, impl, anchor, *tensors):
        ctx.impl = impl
        impl.model._pending_nodes += 1
        out = impl.fwd(ctx, *tensors)
        return out

    @staticmethod
    def backward(ctx, *grads):
        model = ctx.impl.model
        model._backward_begin()
        model._pending_nodes -= 1
        # Last node of this backward pass and it only produces vision-encoder gradients: every other gradient is final,
        # so a data-parallel accelerator may start reducing those ranges now, under the vision backward (the hook may
        # return a callback that the vision backward invokes as each block's gradients become final).
        ctx.impl.block_done = None
        if model._pending_nodes == 0 and getattr(ctx.impl, "kind", None) == "vision" and model._last_node_hook is not None:
            ctx.impl.block_done = model._last_node_hook()
        gin = ctx.impl.bwd(ctx, *grads)
        return (None, None) + tuple(gin)


def _up(g):
    """Upstream gradient of a scalar loss as a contiguous f32 [1] device tensor."""
    return g.reshape(1).to(torch.float32).contiguous()


def _twin(t):
    """bf16 copy of a boundary tensor: the twin the producing op attached, else one cast kernel."""
    tw = getattr(t, "_xfm16", None)
    if tw is not None and tw.shape == t.shape:
        return tw
    src = t.detach().to(torch.float32).contiguous()
    out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    L.cast_to_bf16(src, out)
    return out


def _to_f32(t16):
    out = torch.empty(t16.shape, dtype=torch.float32, device=t16.device)
    L.cast_to_f32(t16.contiguous(), out)
    return out


class _MlpFn(torch.autograd.Function):
    """Linear - LayerNorm - GELU(erf) - Linear on the C-ABI kernels, for heads a task model adds on top of XFMBase."""

    @staticmethod
    def forward(ctx, x, w0, b0, lw, lb, w3, b3):
        x16 = x.detach().to(torch.bfloat16).contiguous()
        w0_16, w3_16 = w0.detach().to(torch.bfloat16), w3.detach().to(torch.bfloat16)
        z = L.gemm(x16, w0_16, bias=b0.detach(), out_dtype=torch.float32)
        zn, stats, _ = L.layernorm_fwd(z, lw.detach(), lb.detach(), 1e-5, out_dtype=torch.float32)
        a = L.gelu_fwd(zn)
        n_out = w3.shape[0]
        pad = (n_out + 7) // 8 * 8
        y = torch.empty((x16.shape[0], pad), dtype=torch.float32, device=x.device)
        L.gemm(a, w3_16, bias=b3.detach(), out=y[:, :n_out])
        ctx.save_for_backward(x16, w0_16, z, zn, stats, a, w3_16, lw.detach())
        ctx.n_out = n_out
        return y[:, :n_out]

    @staticmethod
    def backward(ctx, dy):
        x16, w0_16, z, zn, stats, a, w3_16, lw = ctx.saved_tensors
        R, n_out = dy.shape
        pad = (n_out + 7) // 8 * 8
        dy16 = torch.zeros((R, pad), dtype=torch.bfloat16, device=dy.device)
        dy16[:, :n_out] = dy
        db3 = torch.zeros(pad, dtype=torch.float32, device=dy.device)
        L.colsum_into(dy16, db3)
        dw3 = torch.zeros(w3_16.shape, dtype=torch.float32, device=dy.device)
        BK.wgrad(dw3, dy16[:, :n_out], a)
        da = L.gemm(dy16[:, :n_out], w3_16, b_t=True)
        dzn = L.gelu_bwd(da, zn)
        dlw, dlb = torch.zeros_like(lw), torch.zeros_like(lw)
        dz = L.layernorm_bwd(dzn, z, stats, lw, dlw, dlb, out_dtype=torch.bfloat16)
        db0 = torch.zeros(w0_16.shape[0], dtype=torch.float32, device=dy.device)
        L.colsum_into(dz, db0)
        dw0 = torch.zeros(w0_16.shape, dtype=torch.float32, device=dy.device)
        BK.wgrad(dw0, dz, x16)
        dx = L.gemm(dz, w0_16, b_t=True, out_dtype=torch.float32)
        return dx, dw0, db0, dlw, dlb, dw3, db3[:n_out]


class _Mlp(nn.Sequential):
    """Same parameter names as the reference's nn.Sequential (0.weight, 0.bias, 1.weight, 1.bias, 3.weight, 3.bias)."""

    def forward(self, x):
        shape = x.shape
        y = _MlpFn.apply(x.reshape(-1, shape[-1]), self[0].weight, self[0].bias, self[1].weight, self[1].bias, self[3].weight,
                         self[3].bias)
        return y.reshape(*shape[:-1], y.shape[-1])


def build_mlp(input_dim, output_dim):
    """xfm.py:115-121 — heads built by task models (model_nlvr.py:25 cls_head); forward / backward run on the C-ABI kernels."""
    return _Mlp(nn.Linear(input_dim, input_dim * 2), nn.LayerNorm(input_dim * 2), nn.GELU(), nn.Linear(input_dim * 2, output_dim))


def load_pretrained(model, ckpt_rpath, config, is_eval=False, load_text=False):
    """xfm.py:408-468 for the BEiT-v2 configurations: returns the state_dict to load.  Relative-position tables of another
    resolution are resampled like beit2.py:753-808 (checkpoint.interpolate_rel_pos)."""
    checkpoint = torch.load(ckpt_rpath, map_location="cpu")
    state_dict = checkpoint["model"] if "model" in checkpoint.keys() else checkpoint
    if is_eval:
        return state_dict
    if not config.get("use_beit_v2", True):
        raise ValueError("only the BEiT-v2 vision encoder is built (every shipped config selects it)")
    print("### Loading pretrained vision encoder", flush=True)
    own = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    return CK.finetune_state(state_dict, own, str(config.get("text_encoder", "roberta")), load_text=load_text)


class XFMBase(nn.Module):
    def __init__(self, config=None, load_vision_params=False, load_text_params=False, use_contrastive_loss=False,
                 use_matching_loss=False, use_mlm_loss=False, use_bbox_loss=False, config_text=None, init=None,
                 device=None):
        super().__init__()
        cfg = self.cfg = normalize_config(config)
        config = config or {}
        init = init or default_init()
        self.init_params = []
        self._head_names = {"itm_head": 2, "bbox_head": 4}   # build_mlp heads callable as modules: name -> outputs
        self._head_runners = {}
        fp = self.flat = FlatParams()
        D, Hd = cfg["vision_width"], cfg["hidden"]
        self.vision_width, self.text_width = D, Hd
        self.use_vision_tokenizer = bool(cfg["use_vision_tokenizer"])
        self.detach_text_forMLM = cfg["detach_text_forMLM"]
        self.mim_cls_only = cfg["mim_cls_only"]
        self.text_layers, self.fusion_layers = cfg["text_layers"], cfg["fusion_layers"]
        self.num_text_layers, self.num_cross_layers = cfg["text_layers"], 0

        self.learnable_temp = cfg["learnable_temp"] if use_contrastive_loss else False
        if use_contrastive_loss:
            self.embed_dim = cfg["embed_dim"]
            fp.add("temp", (), init=torch.tensor(float(cfg["temp"])), trainable=self.learnable_temp)
        E.add_vision(fp, cfg, init)
        arch = self.text_arch = cfg["text_arch"]
        E.add_roberta(fp, cfg, init, "text_encoder.", cfg["text_layers"], cross=False, enc_width=D, arch=arch)
        # the fusion encoder is always the RoBERTa class, built from the text encoder's config.json (xfm.py:524-531)
        E.add_roberta(fp, cfg, init, "fusion_encoder.", cfg["fusion_layers"], cross=True, enc_width=D)
        if use_contrastive_loss:
            for n, din in (("vision_proj", D), ("text_proj", Hd)):
                fp.add(n + ".weight", (self.embed_dim, din), init=init(n + ".weight", (self.embed_dim, din)))
                fp.add(n + ".bias", (self.embed_dim,), init=init(n + ".bias", (self.embed_dim,)))
                self.init_params += [n + ".weight", n + ".bias"]
            self.init_params.append("temp")
        self._has_itm, self._has_bbox = use_matching_loss, use_bbox_loss
        for flag, name, nout in ((use_matching_loss, "itm_head", 2), (use_bbox_loss, "bbox_head", 4)):
            if flag:
                E.add_mlp_head(fp, init, name, Hd, nout)
                self.init_params += [f"{name}.{k}" for k in ("0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias")]
        if self.use_vision_tokenizer:
            K, Cd = cfg["codebook_size"], cfg["codebook_dim"]
            fp.add("lm_head.weight", (K, D), init=init("lm_head.weight", (K, D)))
            fp.add("lm_head.bias", (K,), init=init("lm_head.bias", (K,)))
            E.add_vision(fp, cfg, init, prefix="vqkd.encoder.", layerscale=False, relbias=False, abs_pos=True,
                         mask_token=False, trainable=False)
            for n, shape in (("vqkd.encode_task_layer.0.weight", (D, D)), ("vqkd.encode_task_layer.0.bias", (D,)),
                             ("vqkd.encode_task_layer.2.weight", (Cd, D)), ("vqkd.encode_task_layer.2.bias", (Cd,)),
                             ("vqkd.quantize.embedding.weight", (K, Cd))):
                fp.add(n, shape, init=init(n, shape), trainable=False)
        if not self.learnable_temp and use_contrastive_loss:
            self.init_params.remove("temp")
        self._extend_params(fp, cfg, init, config)
        # default: the current CUDA device when there is one, else the host (the reference drivers construct on the host
        # and call .to(device), Pretrain.py:413-417 — see _apply)
        if device is not None:
            dev = torch.device(device)
        else:
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        fp.finalize(dev)
        self.flat_extra, self._extra_params = None, {}

        # ---- module tree with the reference's names
        self._params = {}
        for name, seg in fp.segments.items():
            if name.rsplit(".", 1)[-1].startswith("_"):
                continue
            if name == "temp":
                if not self.learnable_temp:   # xfm.py:505-507: a frozen temp is a python float, not a state_dict entry
                    self.temp = float(cfg["temp"])
                    continue
                p = nn.Parameter(self._temp_alias(), requires_grad=True)
            else:
                p = nn.Parameter(fp.view32(name), requires_grad=seg.trainable)
            self._params[name] = p
            self._register(name, p)
        for enc, a in (("text_encoder.", arch), ("fusion_encoder.", "roberta")):
            # weight tying (xroberta.py:1209-1210,1321-1323; xbert.py:687-692,1536-1537)
            nm = E.TextNames(a)
            h = nm.head("lm_head")
            self._register(enc + h["dec_w"], self._params[enc + nm.stem + "embeddings.word_embeddings.weight"])
            self._register(enc + h["dec_b"], self._params[enc + h["bias"]])
            if a == "roberta":
                self._register(enc + "lm_cap_head.decoder.bias", self._params[enc + "lm_cap_head.bias"])
            self._register_buffer(enc + nm.stem + "embeddings.position_ids", torch.arange(cfg["max_pos"]).expand((1, -1)).clone())
        ws = cfg["image_res"] // cfg["patch_size"]
        rpi = relative_position_index(ws).to(dev)
        for i in range(cfg["vision_depth"]):
            self._register_buffer(f"vision_encoder.blocks.{i}.attn.relative_position_index", rpi)
        # ---- runners
        self._rpi = rpi
        self._vis = E.VisionEncoder(fp, cfg, rel_index=rpi)
        self._txt = E.RobertaStack(fp, cfg, "text_encoder.", cfg["text_layers"], cross=False, arch=arch)
        self._fus = E.RobertaStack(fp, cfg, "fusion_encoder.", cfg["fusion_layers"], cross=True)
        self._vproj = E.ProjHead(fp, "vision_proj") if use_contrastive_loss else None
        self._tproj = E.ProjHead(fp, "text_proj") if use_contrastive_loss else None
        self._itm = E.MlpHead(fp, "itm_head", 2) if use_matching_loss else None
        self._mlm_fus = E.LMHead(fp, cfg, "fusion_encoder.")
        self._mlm_txt = E.LMHead(fp, cfg, "text_encoder.", arch=arch)
        if self.use_vision_tokenizer:
            self._vq = E.VisionEncoder(fp, cfg, prefix="vqkd.encoder.", layerscale=False, relbias=False, abs_pos=True)
            self._mim_head = E.LinearCE(fp, "lm_head", cfg["codebook_size"])
        self._sampler = BlockMaskSampler(ws, cfg["num_masking_patches"], cfg["min_num_patches"])
        self._extend_modules(fp, cfg, config)
        import types
        for enc, nl, fl in (("text_encoder", cfg["text_layers"], cfg["text_layers"]), ("fusion_encoder", cfg["fusion_layers"], 0)):
            # the attributes of RobertaConfig the reference's task models read (model_retrieval.py:16, xfm.py:481-485)
            object.__setattr__(self._modules[enc], "config", types.SimpleNamespace(
                hidden_size=Hd, num_attention_heads=cfg["heads"], num_hidden_layers=nl, fusion_layer=fl, encoder_width=D,
                vocab_size=cfg["vocab_size"], intermediate_size=cfg["ffn"], max_position_embeddings=cfg["max_pos"],
                pad_token_id=cfg["pad_id"], layer_norm_eps=cfg["ln_eps"], type_vocab_size=cfg["type_vocab"]))
        for hn in self._head_names:
            if hn in self._modules:
                object.__setattr__(self._modules[hn], "_owner", lambda s=self: s)
                object.__setattr__(self._modules[hn], "_head_name", hn)
        self._anchor = torch.zeros((), requires_grad=True)
        self._in_backward = False
        self._attached = []
        self._forced_negatives = None   # tests: (image_neg_idx, text_neg_idx) replaces the on-device draw
        self._forced_masks = None       # tests: bool [B, np] replaces the host sampler
        self._static_masks = None       # graph mode: (bool [B, np], int64 [B * n_mask]) device tensors, see graph.py
        self._drop_calls = 0
        self._seed = int(torch.initial_seed()) & 0x7FFFFFFF
        self.last_hard_negative_weights = None
        self._task_ops = None
        self._load_initial_params(config, load_vision_params, load_text_params and use_mlm_loss)
        if not use_mlm_loss:
            assert load_text_params is False  # xfm.py:387-388 (fine-tuning models never load the text checkpoint here)
        named = set(n for n, _ in self.named_parameters())   # xfm.py:517-522
        for n in set(self.init_params):
            if n not in named:
                print(f"warning: {n} not in named_parameters")
                self.init_params.remove(n)

    # ------------------------------------------------------------------ subclass hooks
    def _extend_params(self, fp, cfg, init, config):
        """Task models add their parameters to the flat buffers here (before the buffers are allocated)."""

    def _extend_modules(self, fp, cfg, config):
        """Task models register tied aliases / buffers and build their runners here."""

    # ------------------------------------------------------------------ module plumbing
    def _node(self, path, leaf_cls=_Holder):
        mod = self
        for i, part in enumerate(path):
            if part not in mod._modules:
                cls = _HeadModule if (i == 0 and part in self._head_names) else _Holder
                mod.add_module(part, cls())
            mod = mod._modules[part]
        return mod

    def _register(self, name, param):
        *path, leaf = name.split(".")
        if not path:
            self.register_parameter(leaf, param)
        else:
            self._node(path).register_parameter(leaf, param)

    def _register_buffer(self, name, t):
        *path, leaf = name.split(".")
        self._node(path).register_buffer(leaf, t)

    def _temp_alias(self):
        """0-dim f32 tensor over the flat buffer's `temp` slot with its OWN version counter (not a view of P in autograd
        terms): the reference's `self.temp.clamp_(...)` at the top of every forward (model_pretrain.py:35-37) then does not
        look like an edit of the whole master buffer, which would force a full bf16 shadow recast per step."""
        P = self.flat.P
        off = self.flat.segments["temp"].offset
        return torch.empty(0, dtype=P.dtype, device=P.device).set_(P.untyped_storage(), P.storage_offset() + off, (), ())

    def _temp_dev(self):
        """f32 [1] device view of temp for the kernels (a frozen temp is a python attribute the caller may edit)."""
        t = self.flat.view32("temp").view(1)
        if not self.learnable_temp and getattr(self, "_temp_written", None) != float(self.temp):
            fresh = self.flat._shadow_version == self.flat.P._version
            with torch.no_grad():
                t.fill_(float(self.temp))
            if fresh:
                self.flat._shadow_version = self.flat.P._version
            self._temp_written = float(self.temp)
        return t

    def flat_buffers(self):
        """[(FlatParams, {segment name: nn.Parameter})]: the model's own buffer, then the adopted one (if any)."""
        out = [(self.flat, self._params)]
        if self.flat_extra is not None:
            out.append((self.flat_extra, self._extra_params))
        return out

    def adopt_stray_parameters(self):
        """Trainable parameters created outside the flat buffer — e.g. the reference's own model_nlvr.py:25
        `self.cls_head = build_mlp(...)` assigned after XFMBase.__init__ — move into a second flat buffer (P/G/S), so the
        flat optimizer and the accelerator reduce, clip, update and zero them like every other parameter.  Idempotent."""
        known = {id(p) for p in self._params.values()} | {id(p) for p in self._extra_params.values()}
        stray = [(n, p) for n, p in self.named_parameters() if p.requires_grad and id(p) not in known]
        if not stray:
            return
        if self.flat_extra is not None:
            raise RuntimeError("parameters were added after the optimizer was built: " + ", ".join(n for n, _ in stray[:4]))
        fx = FlatParams()
        for n, p in stray:
            fx.add(n, tuple(p.shape), init=p.detach())
        fx.finalize(self.flat.P.device)
        with torch.no_grad():
            for n, p in stray:
                p.data = fx.view32(n)
                p.grad = fx._view(fx.G, n)
                fx.touched.add(n)
        self.flat_extra, self._extra_params = fx, dict(stray)

    def collect_stray_grads(self):
        """Gradients of adopted parameters accumulate in place in the second buffer; if a driver dropped that .grad
        (set_to_none) autograd allocated a fresh tensor — copy it in and re-attach the view."""
        if self.flat_extra is None:
            return
        fx = self.flat_extra
        for n, p in self._extra_params.items():
            view = fx._view(fx.G, n)
            if p.grad is None:
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view

    def _apply(self, fn, recurse=True):
        """.to(device) / .cuda() move the flat buffers and re-point every parameter view (Pretrain.py:417 `model.to(device)`
        after constructing on the host); dtype changes are refused (fp32 master weights, bf16 compute is internal)."""
        probe = fn(self.flat.P[:1])
        if probe.dtype != torch.float32:
            raise RuntimeError("xfm_b200 keeps fp32 master weights (bf16 compute is internal); .half()/.bfloat16() is not supported")
        known = {id(p) for p in self._params.values()} | {id(p) for p in self._extra_params.values()}
        with torch.no_grad():   # modules a task model attached after construction (model_nlvr.py:25 cls_head)
            for _, p in self.named_parameters():
                if id(p) not in known:
                    p.data = fn(p.data)
                    if p.grad is not None:
                        p.grad.data = fn(p.grad.data)
        if probe.device == self.flat.P.device:
            return self
        for fp, params in self.flat_buffers():
            fp.move(fn)
            with torch.no_grad():
                for name, p in params.items():
                    p.data = self._temp_alias() if (name == "temp" and fp is self.flat) else fp.view32(name)
                    p.grad = None
        if self.flat_extra is not None:
            for n, p in self._extra_params.items():
                p.grad = self.flat_extra._view(self.flat_extra.G, n)
        self._attached = []
        for m in self.modules():
            for k, b in list(m._buffers.items()):
                if b is not None:
                    m._buffers[k] = fn(b)
        dev = self.flat.P.device
        self._rpi = self._rpi.to(dev)
        self._vis.rel_index = self._rpi
        for r in (self._vis, self._txt, self._fus, getattr(self, "_vq", None)) + tuple(self._extra_runners()):
            if r is not None:
                r.w = None
        self._task_ops = None
        self._temp_written = None
        return self

    def _extra_runners(self):
        """Task models list the runners (objects caching parameter views in `.w`) they built in _extend_modules."""
        return ()

    def _load_initial_params(self, config, load_vision_params, load_text_params):
        """Constructor-time checkpoint import of the reference (xfm.py:205-256 vision, :298-385 text, model_vqkd.py:315-333
        tokenizer).  Adds the text-encoder keys the checkpoint lacks to init_params (xfm.py:385)."""
        config = config or {}
        own = None
        if self.use_vision_tokenizer and config.get("tokenizer_weight") and os.path.exists(str(config["tokenizer_weight"])):
            self.load_state_dict(CK.vqkd_state(config["tokenizer_weight"]), strict=False)
        if load_vision_params:
            from .config import _read_json
            vis = _read_json(config.get("vision_config")) or {}
            own = {k: tuple(v.shape) for k, v in self.state_dict().items()}
            sd = CK.beit2_init_state(vis["ckpt"], own, self.cfg["vision_depth"])
            msg = self.load_state_dict(sd, strict=False)
            missing = [k for k in msg.missing_keys if k.startswith("vision_encoder.") and "relative_position_index" not in k]
            if missing:
                print("Weights of VisionTransformer not initialized from pretrained model: {}".format(missing))
            if msg.unexpected_keys:
                print("Weights from pretrained model not used in VisionTransformer: {}".format(msg.unexpected_keys))
        if load_text_params:
            sd = CK.text_init_state(str(config["text_encoder"]), self.cfg["text_layers"], self.text_arch)
            msg = self.load_state_dict(sd, strict=False)
            missing = [k[len("text_encoder."):] for k in msg.missing_keys if k.startswith("text_encoder.")]
            print("missing_keys: ", missing, flush=True)
            print("unexpected_keys: ", [k[len("text_encoder."):] for k in msg.unexpected_keys], flush=True)
            self.init_params += [f"text_encoder.{k}" for k in missing]

    def load_state_dict(self, state_dict, strict=True, assign=False):
        # fine-tuning models of the reference hold a bare RobertaModel as text_encoder (xfm.py:397-403), so their
        # checkpoints name it text_encoder.<x>; this module always keeps the pre-training layout text_encoder.roberta.<x>
        stem = "text_encoder." + self.text_arch + "."
        state_dict = {((stem + k[len("text_encoder."):])
                       if k.startswith(("text_encoder.embeddings.", "text_encoder.encoder.")) else k): v
                      for k, v in state_dict.items()}
        out = super().load_state_dict(state_dict, strict=strict, assign=False)
        self.flat.sync_shadow(force=True)   # (on the host: only marks the shadow stale, Pretrain.py:413-417)
        return out

    def load_pretrained(self, ckpt_rpath, config, is_eval=False, is_domain_pretrain=False):
        """xfm.py:542-557: a domain-pre-training checkpoint is loaded as it is ('visual_encoder' keys renamed); anything else goes
        through the module-level load_pretrained (vision tables resampled to this model's resolution, text keys renamed)."""
        if is_domain_pretrain:
            ck = torch.load(ckpt_rpath, map_location="cpu")
            sd = ck["model"] if "model" in ck else ck
            sd = {k.replace("visual_encoder", "vision_encoder"): v for k, v in sd.items()}
        else:
            sd = load_pretrained(self, ckpt_rpath, config, is_eval=is_eval, load_text=True)
        msg = self.load_state_dict(sd, strict=False)
        print("load checkpoint from %s" % ckpt_rpath)
        print("missing_keys: ", [p for p in msg.missing_keys if "vision_encoder" not in p])
        print("unexpected_keys: ", msg.unexpected_keys)

    # ------------------------------------------------------------------ backward bookkeeping
    def _backward_begin(self):
        if self._in_backward:
            return
        self._in_backward = True
        if self._backward_begin_hook is not None:
            self._backward_begin_hook()
        if self._attached and self._attached[0].grad is None:  # optimizer.zero_grad(set_to_none=True) ran
            self.flat.zero_grad()
            self._attached = []
        torch.autograd.Variable._execution_engine.queue_callback(self._backward_end)

    def _backward_end(self):
        self._in_backward = False
        self._pending_nodes = 0
        G = self.flat
        att = []
        for name in G.touched:
            p = self._params.get(name)
            if p is not None and p.requires_grad:
                if p.grad is None:
                    p.grad = G._view(G.G, name)
                att.append(p)
        self._attached = att

    def zero_grad(self, set_to_none=True):
        self.flat.zero_grad()
        for p in self._params.values():
            p.grad = None
        self._attached = []
        if self.flat_extra is not None:   # adopted parameters keep their .grad views (autograd accumulates in place)
            self.flat_extra.G.zero_()
            for n, p in self._extra_params.items():
                p.grad = self.flat_extra._view(self.flat_extra.G, n)

    def _prep(self):
        self.flat.sync_shadow()

    def clamp_temp(self, lo, hi):
        """temp.clamp_(lo, hi) (model_pretrain.py:35-37).  `temp` has its own version counter (_temp_alias), so neither this
        nor the reference's own in-place clamp invalidates the bf16 weight shadow (temp is only ever read in fp32)."""
        with torch.no_grad():
            self._params["temp"].clamp_(lo, hi)

    def _drop(self):
        if not self.training:
            return BK.NO_DROP
        self._drop_calls += 1
        return BK.DropCfg(self.cfg["hidden_dropout"], self.cfg["attn_dropout"], self._seed * 131 + self._drop_calls)

    _pending_nodes = 0
    _masked_vision = None   # (image, (embeds, atts, mask)) of a twin vision pass, for the do_mask call that follows
    _last_node_hook = None
    _backward_begin_hook = None

    def _call(self, impl, *tensors):
        impl.model = self
        if torch.is_grad_enabled():
            impl.save = True
            return _Op.apply(impl, self._anchor, *tensors)
        impl.save = False
        return impl.fwd(None, *tensors)

    # ------------------------------------------------------------------ vision
    def get_vision_embeds(self, image, image_atts=None, idx_to_group_img=None, do_mask=False, _also_masked=False):
        """xfm.py:560-597.  With idx_to_group_img (region data: fewer images than samples) the encoded images are gathered per
        sample; with image_atts as well, token 0 becomes the region-weighted mean of the sample's patch tokens
        (beit2.py:468-475) and the full-image embeddings are returned as a third value.
        _also_masked (xfm_b200 only): the caller will ask for the masked copy of the SAME images next
        (get_vision_embeds(image, do_mask=True), model_pretrain.py:73).  Both copies then run as one 2B-sample pass of the
        encoder (VisionEncoder.forward twin=True); the masked half is kept for that next call.  Masks come from the same
        sampler calls in the same order, so they are the ones the separate pass would have drawn."""
        if idx_to_group_img is not None:
            return self._region_vision_embeds(image, image_atts, idx_to_group_img)
        kept = self._masked_vision
        self._masked_vision = None
        if do_mask and kept is not None and kept[0] is image:
            return kept[1]
        self._prep()
        B = image.shape[0]
        mask_dev = None
        want_mask = do_mask or _also_masked
        if want_mask and self._static_masks is not None:   # CUDA-graph mode: device buffers refilled between replays
            mask_dev, rows = self._static_masks
        elif want_mask:
            if self._forced_masks is not None:
                m = self._forced_masks.cpu()
                rows = torch.from_numpy(__import__("numpy").flatnonzero(m.numpy()).astype("int64"))
            else:
                m, rows = sample_batch(self._sampler, B)
            # pinned staging: a pageable source would make the "non_blocking" copy synchronise with the device
            mask_dev = m.pin_memory().to(image.device, non_blocking=True)
            rows = rows.pin_memory().to(image.device, non_blocking=True)
        model = self
        twin = bool(_also_masked) and not do_mask

        class Impl:
            def fwd(self, ctx, image):
                mu8 = mask_dev.to(torch.uint8) if mask_dev is not None else None
                y32, y16, st = model._vis.forward(image, mask_u8=mu8, train=model.training, save=self.save, twin=twin)
                self.y16 = y16
                if ctx is not None:
                    ctx.st = st
                if twin:
                    return y32[:B], y32[B:]
                return y32

            def bwd(self, ctx, *grads):
                if twin:   # one backward pass over both copies, once both gradients are there
                    z = [g if g is not None else torch.zeros((B,) + tuple(self.y16.shape[1:]), dtype=torch.float32,
                                                             device=self.y16.device) for g in grads]
                    dy = torch.cat([z[0].float(), z[1].float()])
                else:
                    dy = grads[0]
                model._vis.backward(ctx.st, dy, block_done=getattr(self, "block_done", None))
                ctx.st = None
                return (None,)
        impl = Impl()
        impl.kind = "vision"
        y = self._call(impl, image)

        def atts_of(t):
            a = torch.ones(t.shape[:-1], dtype=torch.long, device=image.device)
            a._xfm_all_ones = True   # lets the cross-attention skip the encoder mask (and stay on the tcgen05 kernel)
            return a
        if twin:
            y, ym = y
            y._xfm16, ym._xfm16 = impl.y16[:B], impl.y16[B:]
            mask_dev._xfm_rows = rows
            self._masked_vision = (image, (ym, atts_of(ym), mask_dev))
            return y, atts_of(y)
        y._xfm16 = impl.y16
        if do_mask:
            mask_dev._xfm_rows = rows
            return y, atts_of(y), mask_dev
        return y, atts_of(y)

    def _region_vision_embeds(self, image, image_atts, idx_to_group_img):
        """xfm.py:574-597 + beit2.py:468-475."""
        y, _ = self.get_vision_embeds(image)
        model = self
        idx = idx_to_group_img.reshape(-1).long().contiguous()
        n_img, N, D = y.shape
        bsz = idx.numel()
        y16 = _twin(y)
        pooled = image_atts is not None
        if pooled:
            assert image_atts.size(0) == idx.size(0)
            atts = image_atts.long().contiguous()

        class Impl:
            def fwd(self, ctx, y_):
                y32 = y_.detach().float().contiguous()
                full32 = L.gather_rows(y32.view(n_img, N * D), idx).view(bsz, N, D)
                self.full16 = L.gather_rows(y16.reshape(n_img, N * D), idx).view(bsz, N, D)
                if not pooled:
                    return (full32,)
                out32, self.out16 = L.region_pool_fwd(y32, idx, atts)
                return out32, full32

            def bwd(self, ctx, *grads):
                dy = torch.zeros((n_img, N * D), dtype=torch.float32, device=y.device)
                d_full = grads[-1]
                if d_full is not None:
                    L.scatter_add_rows_(dy, idx, d_full.float().reshape(bsz, N * D).contiguous())
                if pooled and grads[0] is not None:
                    L.region_pool_bwd_(grads[0].float().contiguous(), idx, atts, dy.view(n_img, N, D))
                return (dy.view(n_img, N, D),)
        impl = Impl()
        outs = self._call(impl, y)
        full = outs[-1]
        full._xfm16 = impl.full16
        if not pooled:
            ones = torch.ones(full.shape[:-1], dtype=torch.long, device=image.device)
            ones._xfm_all_ones = True
            return full, ones
        emb = outs[0]
        emb._xfm16 = impl.out16
        return emb, image_atts, full

    @staticmethod
    def _enc_kmask(image_atts, index=None):
        """Additive key mask of the cross-attention, f32 [samples, image tokens]: (1 - m) * -1e9 (HF invert_attention_mask
        for fp32, xroberta.py:903-909), or None for the all-ones masks get_vision_embeds produces.  index: sample -> mask row."""
        if image_atts is None or getattr(image_atts, "_xfm_all_ones", False):
            return None
        if getattr(getattr(image_atts, "_base", None), "_xfm_all_ones", False):
            return None   # a slice view of get_vision_embeds' all-ones mask (model_nlvr.py:33-36 splits it per image)
        m = ((1.0 - image_atts.to(torch.float32)) * -1e9)
        if index is not None:
            m = m.index_select(0, index.long())
        return m.contiguous()

    # ------------------------------------------------------------------ text
    def get_text_embeds(self, text_ids, text_atts, _also_masked=None):
        """xfm.py:600-611: 12-layer text encoder (no cross-attention).
        _also_masked (xfm_b200 only): the masked copy of the same texts, which get_fuse_mlm_loss would encode in a second,
        gradient-free pass (xfm.py:645-649, detach_text_forMLM).  Both copies then run as ONE 2B-sample pass — the 40-token
        GEMMs of a B-sample pass sit on the launch-latency floor, a 2B-sample pass costs barely more — and the masked half is
        handed to get_matching_and_fuse_mlm_loss; the backward runs over the clean half only."""
        assert text_atts is not None
        self._prep()
        model = self
        B, Lt = text_ids.shape
        self._masked_text = None
        pair = _also_masked is not None and self.detach_text_forMLM
        ids_all = torch.cat([text_ids, _also_masked]) if pair else text_ids
        Ball = ids_all.shape[0]
        kmask1 = E.RobertaStack.additive_mask(text_atts)
        kmask = torch.cat([kmask1, kmask1]) if pair else kmask1

        class Impl:
            def fwd(self, ctx):
                drop = model._drop()
                h, h32, est = model._txt.embed(ids_all, drop, save=self.save)
                h, h32, st = model._txt.layers_fwd(h, Ball, Lt, kmask, drop=drop, save=self.save, h32=h32)
                if pair:
                    model._masked_text = (_also_masked, h[B * Lt:], h32[B * Lt:])
                    h, h32 = h[:B * Lt], h32[:B * Lt]
                    if self.save:
                        est, st = E.RobertaStack.slice_states(est, st, B)
                self.h16 = h.view(B, Lt, -1)
                if ctx is not None:
                    ctx.est, ctx.st = est, st
                return h32.view(B, Lt, -1)

            def bwd(self, ctx, dh):
                d = model._txt.layers_bwd(ctx.st, dh.reshape(B * Lt, -1).contiguous(), need_dh=True)
                model._txt.embed_bwd(ctx.est, d)
                ctx.st = ctx.est = None
                return ()
        impl = Impl()
        y = self._call(impl)
        y._xfm16 = impl.h16
        return y

    # ------------------------------------------------------------------ ITC features
    def get_features(self, image_embeds=None, text_embeds=None):
        """xfm.py:614-621."""
        self._prep()
        model = self

        def one(head, emb):
            B, Lt, _ = emb.shape

            class Impl:
                def fwd(self, ctx, emb_):
                    y, st = head.forward(emb_.detach().to(torch.float32).contiguous(), B, Lt)
                    if ctx is not None:
                        ctx.st = st
                    return y

                def bwd(self, ctx, dy):
                    return (head.backward(ctx.st, dy),)
            return model._call(Impl(), emb)
        if image_embeds is None:
            return one(self._tproj, text_embeds)
        if text_embeds is None:
            return one(self._vproj, image_embeds)
        return one(self._vproj, image_embeds), one(self._tproj, text_embeds)

    # ------------------------------------------------------------------ ITC loss
    def _gather_world(self, image_feat, text_feat, idx):
        """AllGather of xfm.py:81-101: every rank's features in rank order; backward keeps the local slice."""
        dist = torch.distributed
        B, Ed = image_feat.shape
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return image_feat.contiguous(), text_feat.contiguous(), idx, 0
        W, r = dist.get_world_size(), dist.get_rank()
        msg = torch.cat([image_feat, text_feat], 1).contiguous()  # one message per rank: [B, 2E]
        allm = torch.empty((W * B, 2 * Ed), dtype=msg.dtype, device=msg.device)  # rank-major concatenation along dim 0
        dist.all_gather_into_tensor(allm, msg)
        idx_all = None
        if idx is not None:
            idx_all = torch.empty((W * B,), dtype=idx.dtype, device=idx.device)
            dist.all_gather_into_tensor(idx_all, idx.reshape(-1).contiguous())
        return allm[:, :Ed].contiguous(), allm[:, Ed:].contiguous(), idx_all, r * B

    def get_contrastive_loss(self, image_feat, text_feat, idx=None):
        """xfm.py:683-715: loss and (in the same kernel sequence) its gradient wrt the local features and temp."""
        assert image_feat.size(-1) == self.embed_dim
        assert text_feat.size(-1) == self.embed_dim
        model = self
        B = image_feat.shape[0]
        if idx is not None:
            idx = idx.view(-1)
            assert idx.size(0) == B
        temp = self._temp_dev()

        class Impl:
            def fwd(self, ctx, fi, ft):
                ia, ta, idx_all, off = model._gather_world(fi.detach().float(), ft.detach().float(), idx)
                loss, di, dt, dtemp = L.itc_loss_fused(ia, ta, temp, off, B, idx_all)
                if ctx is not None:
                    ctx.g = (di, dt, dtemp)
                return loss.view(())

            def bwd(self, ctx, g):
                di, dt, dtemp = ctx.g
                up = _up(g)
                if model.learnable_temp:
                    model.flat.grad("temp").add_((dtemp * up).view(()))
                return L.scale_by_scalar(di, up), L.scale_by_scalar(dt, up)
        return self._call(Impl(), image_feat, text_feat)

    # ------------------------------------------------------------------ ITM
    def get_hard_negatives(self, image_feat, text_feat, idx=None):
        """xfm.py:717-746 with the multinomial draw on the device: returns (image_neg_idx, text_neg_idx) int64 [B]."""
        if self._forced_negatives is not None:
            i, t = self._forced_negatives
            return i.to(image_feat.device), t.to(image_feat.device)
        self._drop_calls += 1
        seed = (self._seed * 7919 + self._drop_calls * 104729) & 0x7FFFFFFFFFFFFFFF
        ineg, tneg, _, _ = L.hard_negatives(image_feat.detach().float().contiguous(), text_feat.detach().float().contiguous(),
                                            self._temp_dev(), seed, idx=None if idx is None else idx.view(-1))
        return ineg, tneg

    def hard_negative_weights(self, image_feat, text_feat, idx=None):
        """The deterministic part of get_hard_negatives (weights_i2t, weights_t2i), for parity checks."""
        _, _, w1, w2 = L.hard_negatives(image_feat.detach().float().contiguous(), text_feat.detach().float().contiguous(),
                                        self._temp_dev(), 0, idx=None if idx is None else idx.view(-1),
                                        want_weights=True)
        return w1, w2

    def _fusion_run(self, text16, Bt, Lt, kmask, img16, Bi, kv_index, save, text32=None, enc_kmask=None):
        """fusion encoder over Bt text samples attending to Bi images (kv_index: sample -> image row, or None).
        enc_kmask: additive key mask over the image tokens per text sample (region data), else all tokens are visible."""
        Ni = img16.shape[1]
        enc = img16.reshape(Bi * Ni, -1)
        if text32 is not None:
            text32 = text32.reshape(Bt * Lt, -1)
        return self._fus.layers_fwd(text16.reshape(Bt * Lt, -1), Bt, Lt, kmask, enc=enc, Benc=Bi, Lenc=Ni, kv_index=kv_index,
                                    drop=self._drop(), save=save, h32=text32, enc_kmask=enc_kmask)

    def _fusion_back(self, st, dh, Bi, Ni, need_dtext, kv_index):
        # with the layers' K | V projections merged, ONE GEMM writes the image-token gradient: no zero fill, no accumulate
        fresh = getattr(st, "kvc_all", None) is not None
        alloc = torch.empty if fresh else torch.zeros
        d_enc = alloc((Bi * Ni, self.vision_width), dtype=torch.float32, device=dh.device)
        d_text = self._fus.layers_bwd(st, dh, d_enc=d_enc, need_dh=need_dtext, d_enc_fresh=fresh)  # CSR: built in layers_fwd
        return d_text, d_enc.view(Bi, Ni, -1)

    def get_cross_embeds(self, image_embeds, image_atts, text_ids=None, text_embeds=None, text_atts=None, is_pretrain=True):
        """xfm.py:659-680.  image_atts other than get_vision_embeds' all-ones mask (region data) become an additive key mask of
        the cross-attention."""
        self._prep()
        model = self
        Bi, Ni, _ = image_embeds.shape
        kmask = E.RobertaStack.additive_mask(text_atts)
        ek = self._enc_kmask(image_atts)
        img16 = _twin(image_embeds)
        Bt, Lt = text_atts.shape
        assert Bt == Bi
        from_ids = text_embeds is None
        need_dtext = (not from_ids) and (not is_pretrain)
        t16 = None if from_ids else _twin(text_embeds)

        class Impl:
            def fwd(self, ctx, img, txt):
                est = None
                h, h32 = t16, (None if from_ids else txt.detach().float().contiguous())
                if from_ids:
                    h, h32, est = model._fus.embed(text_ids, model._drop(), save=self.save)
                out, out32, st = model._fusion_run(h, Bt, Lt, kmask, img16, Bi, None, self.save, text32=h32, enc_kmask=ek)
                self.h16 = out.view(Bt, Lt, -1)
                if ctx is not None:
                    ctx.st, ctx.est = st, est
                return out32.view(Bt, Lt, -1)

            def bwd(self, ctx, dh):
                d_text, d_img = model._fusion_back(ctx.st, dh.reshape(Bt * Lt, -1).contiguous(), Bi, Ni,
                                                   need_dtext or from_ids, None)
                if from_ids:
                    model._fus.embed_bwd(ctx.est, d_text)
                    d_text = None
                ctx.st = None
                return d_img, (d_text.float().view(Bt, Lt, -1) if (d_text is not None and need_dtext) else None)
        impl = Impl()
        dummy = text_embeds if text_embeds is not None else image_embeds.new_zeros(())
        y = self._call(impl, image_embeds, dummy)
        y._xfm16 = impl.h16
        return y

    def _head_apply(self, name, x):
        """model.<head>(x) for a build_mlp head living in the flat buffer: x f32/bf16 [..., din] -> logits f32 [..., nout]."""
        self._prep()
        model = self
        runner = self._head_runners.get(name)
        if runner is None:
            runner = self._head_runners[name] = E.MlpHead(self.flat, name, self._head_names[name])
        shape = x.shape
        nout = self._head_names[name]

        class Impl:
            def fwd(self, ctx, x_):
                x2 = x_.detach().reshape(-1, shape[-1])
                x16 = (x2 if x2.dtype == torch.bfloat16 else _twin(x2)).contiguous()
                logits, st = runner.logits(x16, save=self.save)
                if ctx is not None:
                    ctx.st = st
                return logits.reshape(*shape[:-1], nout)

            def bwd(self, ctx, dlog):
                R = ctx.st.x0.shape[0]
                d16 = torch.zeros((R, 8), dtype=torch.bfloat16, device=dlog.device)
                d16[:, :nout] = dlog.reshape(R, nout)
                dx = torch.empty((R, shape[-1]), dtype=torch.bfloat16, device=dlog.device)
                runner.backward(ctx.st, d16, dx)
                ctx.st = None
                return (dx.to(x.dtype).reshape(shape),)
        return self._call(Impl(), x)

    def get_matching_loss(self, image_embeds, image_atts, image_feat, text_ids, text_atts, text_feat, idx=None,
                          return_cross_embeds=False, text_embeds=None, is_pretrain=True):
        """xfm.py:749-802.  The B positive and 2B hard-negative pairs run as ONE 3B-sample fusion pass in which every
        sample indexes the image whose K/V it attends to (negative images are rows of the same batch), so the
        cross-attention K/V projection is computed once per image instead of three times."""
        assert text_ids.dim() == 2, "X-Brain uses text_ids for matching."
        assert text_embeds is not None, "xfm_b200 matches on text_embeds (the only call pattern of the reference drivers)"
        self._prep()
        model = self
        B, Ni, _ = image_embeds.shape
        Lt = text_atts.shape[1]
        with torch.no_grad():
            image_neg, text_neg = self.get_hard_negatives(image_feat, text_feat, idx=idx)
        ar = torch.arange(B, device=image_embeds.device)
        # sample order = reference order: pos [B] | (image_neg, text) [B] | (image, text_neg) [B]   (xfm.py:782-797)
        txt_index = torch.cat([ar, ar, text_neg.long()])
        kv_index = torch.cat([ar, image_neg.long(), ar]).to(torch.int32)
        kmask = E.RobertaStack.additive_mask(text_atts.index_select(0, txt_index))
        ek = self._enc_kmask(image_atts, kv_index)
        img16, t16 = _twin(image_embeds), _twin(text_embeds)
        need_dtext = not is_pretrain
        labels = torch.zeros(3 * B, dtype=torch.long, device=image_embeds.device)
        labels[:B] = 1

        class Impl:
            def fwd(self, ctx, img, txt):
                D = t16.shape[-1]
                tall = L.gather_rows(t16.reshape(B, Lt * D), txt_index).view(3 * B * Lt, D)
                tall32 = L.gather_rows(txt.detach().float().reshape(B, Lt * D).contiguous(), txt_index)
                h, _, st = model._fusion_run(tall, 3 * B, Lt, kmask, img16, B, kv_index, self.save, text32=tall32, enc_kmask=ek)
                x0 = E.cls_rows(h, 3 * B, Lt)
                logits, hst = model._itm.logits(x0, save=self.save)
                loss, count, lse = L.ce_fwd(logits, labels, 2)
                self.cross_pos = x0[:B].float()
                if ctx is not None:
                    ctx.st, ctx.hst, ctx.ce = st, hst, (logits, count, lse)
                return loss.view(())

            def bwd(self, ctx, g):
                logits, count, lse = ctx.ce
                dlog = L.ce_bwd(logits, labels, lse, count, _up(g), 2, 8)
                D = t16.shape[-1]
                dh = torch.zeros((3 * B * Lt, D), dtype=torch.bfloat16, device=dlog.device)
                model._itm.backward(ctx.hst, dlog, E.cls_rows(dh, 3 * B, Lt))
                d_tall, d_img = model._fusion_back(ctx.st, dh, B, Ni, need_dtext, kv_index)
                ctx.st = None
                d_txt = None
                if need_dtext:
                    d_txt = torch.zeros((B, Lt * D), dtype=torch.float32, device=dlog.device)
                    L.scatter_add_rows_(d_txt, txt_index, d_tall.reshape(3 * B, Lt * D).contiguous())
                    d_txt = d_txt.view(B, Lt, D)
                return d_img, d_txt
        impl = Impl()
        loss = self._call(impl, image_embeds, text_embeds)
        if return_cross_embeds:
            return loss, impl.cross_pos
        return loss

    def get_matching_and_fuse_mlm_loss(self, image_embeds, image_atts, image_feat, text_ids, text_atts, text_feat,
                                       text_ids_masked, masked_pos, masked_ids, idx=None, text_embeds=None, is_pretrain=True):
        """get_matching_loss (xfm.py:749-802) and get_fuse_mlm_loss (xfm.py:638-656) as ONE 4B-sample pass of the fusion
        encoder: rows [0,B) positives, [B,2B) (image_neg, text), [2B,3B) (image, text_neg), [3B,4B) the masked-LM texts.  All
        four groups attend to the same B images, so every layer's cross-attention K/V projection (and its backward) runs
        once instead of twice and the text-side GEMMs see M = 4·B·L rows.  Per-sample arithmetic is unchanged; returns
        (loss_itm, loss_mlm) equal to the two separate calls (tests/test_model_gpu.py)."""
        assert text_embeds is not None and text_ids.dim() == 2
        self._prep()
        model = self
        B, Ni, _ = image_embeds.shape
        Lt = text_atts.shape[1]
        dev = image_embeds.device
        with torch.no_grad():
            image_neg, text_neg = self.get_hard_negatives(image_feat, text_feat, idx=idx)
        ar = torch.arange(B, device=dev)
        txt_index = torch.cat([ar, ar, text_neg.long()])
        kv_index = torch.cat([ar, image_neg.long(), ar, ar]).to(torch.int32)
        kmask_mlm = E.RobertaStack.additive_mask(text_atts)
        kmask = torch.cat([kmask_mlm.index_select(0, txt_index), kmask_mlm]).contiguous()
        ek = self._enc_kmask(image_atts, kv_index)
        img16, t16 = _twin(image_embeds), _twin(text_embeds)
        need_dtext = not is_pretrain
        detach = self.detach_text_forMLM
        itm_labels = torch.zeros(3 * B, dtype=torch.long, device=dev)
        itm_labels[:B] = 1
        rows = (torch.arange(B, device=dev).view(B, 1) * Lt + masked_pos).reshape(-1).contiguous() + 3 * B * Lt
        mlm_labels = masked_ids.reshape(-1).contiguous()
        head = self._mlm_fus

        class Impl:
            def fwd(self, ctx, img, txt):
                D = t16.shape[-1]
                drop = model._drop()
                keep_text = self.save and not detach
                cached, model._masked_text = model._masked_text, None
                if cached is not None and cached[0] is text_ids_masked and detach:   # encoded together with the clean texts
                    hm, hm32, est, tst = cached[1], cached[2], None, None
                else:
                    hm, hm32, est = model._txt.embed(text_ids_masked, drop, save=keep_text)
                    hm, hm32, tst = model._txt.layers_fwd(hm, B, Lt, kmask_mlm, drop=drop, save=keep_text, h32=hm32)
                tall = torch.empty((4 * B * Lt, D), dtype=torch.bfloat16, device=dev)
                tall32 = torch.empty((4 * B * Lt, D), dtype=torch.float32, device=dev)
                tall[:3 * B * Lt] = L.gather_rows(t16.reshape(B, Lt * D), txt_index).view(3 * B * Lt, D)
                tall32[:3 * B * Lt] = L.gather_rows(txt.detach().float().reshape(B, Lt * D).contiguous(), txt_index).view(3 * B * Lt, D)
                tall[3 * B * Lt:] = hm
                tall32[3 * B * Lt:] = hm32
                h, _, st = model._fusion_run(tall, 4 * B, Lt, kmask, img16, B, kv_index, self.save, text32=tall32, enc_kmask=ek)
                x0 = E.cls_rows(h, 4 * B, Lt)[:3 * B]
                logits, hst = model._itm.logits(x0, save=self.save)
                loss_itm, count, lse = L.ce_fwd(logits, itm_labels, 2)
                x = L.gather_rows(h, rows)
                loss_mlm, mst = head.loss(x, mlm_labels)
                if ctx is not None:
                    ctx.st, ctx.hst, ctx.ce, ctx.mst, ctx.est, ctx.tst = st, hst, (logits, count, lse), mst, est, tst
                return loss_itm.view(()), loss_mlm.view(())

            def bwd(self, ctx, g_itm, g_mlm):
                D = t16.shape[-1]
                dh = torch.zeros((4 * B * Lt, D), dtype=torch.float32, device=dev)
                if g_itm is not None:
                    logits, count, lse = ctx.ce
                    dlog = L.ce_bwd(logits, itm_labels, lse, count, _up(g_itm), 2, 8)
                    model._itm.backward(ctx.hst, dlog, E.cls_rows(dh, 4 * B, Lt)[:3 * B])
                if g_mlm is not None:
                    dx = head.backward(ctx.mst, _up(g_mlm))
                    L.scatter_add_rows_(dh, rows, dx)
                need_any = need_dtext or not detach
                d_tall, d_img = model._fusion_back(ctx.st, dh, B, Ni, need_any, kv_index)
                d_txt = None
                if need_dtext:
                    d_txt = torch.zeros((B, Lt * D), dtype=torch.float32, device=dev)
                    L.scatter_add_rows_(d_txt, txt_index, d_tall[:3 * B * Lt].reshape(3 * B, Lt * D).contiguous())
                    d_txt = d_txt.view(B, Lt, D)
                if not detach:
                    d = model._txt.layers_bwd(ctx.tst, d_tall[3 * B * Lt:].contiguous(), need_dh=True)
                    model._txt.embed_bwd(ctx.est, d)
                ctx.st = ctx.hst = ctx.ce = ctx.mst = ctx.est = ctx.tst = None
                return d_img, d_txt
        impl = Impl()
        return self._call(impl, image_embeds, text_embeds)

    # ------------------------------------------------------------------ MLM
    def _mlm(self, text_ids_masked, text_atts, image_embeds, masked_pos, masked_ids, fused, image_atts=None):
        self._prep()
        model = self
        B, Lt = text_ids_masked.shape
        M = masked_pos.shape[1]
        kmask = E.RobertaStack.additive_mask(text_atts)
        rows = (torch.arange(B, device=masked_pos.device).view(B, 1) * Lt + masked_pos).reshape(-1).contiguous()
        labels = masked_ids.reshape(-1).contiguous()
        img16 = _twin(image_embeds) if fused else None
        ek = self._enc_kmask(image_atts) if fused else None
        detach = fused and self.detach_text_forMLM
        head = self._mlm_fus if fused else self._mlm_txt

        class Impl:
            def fwd(self, ctx, img):
                drop = model._drop()
                keep_text = self.save and not detach
                h, h32, est = model._txt.embed(text_ids_masked, drop, save=keep_text)
                h, h32, tst = model._txt.layers_fwd(h, B, Lt, kmask, drop=drop, save=keep_text, h32=h32)
                fst = None
                if fused:
                    h, h32, fst = model._fusion_run(h, B, Lt, kmask, img16, img16.shape[0], None, self.save, text32=h32,
                                                    enc_kmask=ek)
                x = L.gather_rows(h, rows)
                loss, hst = head.loss(x, labels)
                if ctx is not None:
                    ctx.est, ctx.tst, ctx.fst, ctx.hst = est, tst, fst, hst
                return loss.view(())

            def bwd(self, ctx, g):
                dx = head.backward(ctx.hst, _up(g))
                D = dx.shape[1]
                dh = torch.zeros((B * Lt, D), dtype=torch.float32, device=dx.device)
                L.scatter_add_rows_(dh, rows, dx)
                d_img = None
                if fused:
                    dh, d_img = model._fusion_back(ctx.fst, dh, img16.shape[0], img16.shape[1], not detach, None)
                if not detach:
                    d = model._txt.layers_bwd(ctx.tst, dh, need_dh=True)
                    model._txt.embed_bwd(ctx.est, d)
                ctx.est = ctx.tst = ctx.fst = ctx.hst = None
                return (d_img,)
        dummy = image_embeds if fused else self._anchor
        return self._call(Impl(), dummy)

    def get_fuse_mlm_loss(self, text_ids_masked, text_atts, image_embeds, image_atts, masked_pos, masked_ids):
        """xfm.py:638-656: text-encode the masked ids (detached), fuse with the image, LM head + CE on masked positions."""
        return self._mlm(text_ids_masked, text_atts, image_embeds, masked_pos, masked_ids, fused=True, image_atts=image_atts)

    def get_mlm_loss(self, text_ids_masked, text_atts, image_embeds, image_atts, masked_pos, masked_ids):
        """xfm.py:805-812 on the text-only stream (model_pretrain.py:93-98 passes image_embeds=None)."""
        if image_embeds is not None:
            raise NotImplementedError("text_encoder has no cross-attention in the shipped layout; use get_fuse_mlm_loss")
        return self._mlm(text_ids_masked, text_atts, None, masked_pos, masked_ids, fused=False)

    # ------------------------------------------------------------------ boxes (region data)
    def predict_bbox(self, image_embeds, text_ids, text_atts, text_embeds, is_pretrain=True):
        """xfm.py:843-854: fusion over the FULL image tokens, bbox_head on the CLS row, sigmoid -> (cx, cy, w, h) [bsz, 4]."""
        assert image_embeds.size(0) == text_ids.size(0) == text_atts.size(0)
        ones = torch.ones(image_embeds.shape[:2], dtype=torch.long, device=image_embeds.device)
        ones._xfm_all_ones = True
        output_cls = self.get_cross_embeds(image_embeds, ones, text_ids=text_ids, text_atts=text_atts, text_embeds=text_embeds,
                                           is_pretrain=is_pretrain)[:, 0, :]
        logits = self.bbox_head(output_cls)
        model = self

        class Impl:
            def fwd(self, ctx, x):
                y = L.sigmoid_fwd(x.detach().float().contiguous())
                if ctx is not None:
                    ctx.y = y
                return y

            def bwd(self, ctx, dy):
                return (L.sigmoid_bwd(dy.float().contiguous(), ctx.y),)
        return self._call(Impl(), logits)

    def get_bbox_loss(self, output_coord, target_bbox, is_image=None):
        """xfm.py:815-840: L1 and generalized-IoU losses over the region samples (is_image = 1 rows excluded), one kernel for
        both losses and their gradients; the degenerate-box early-out (xfm.py:825-828) is decided on the device."""
        model = self
        tgt = target_bbox.detach().float().contiguous()
        keep = None if is_image is None else is_image.detach().float().contiguous()

        class Impl:
            def fwd(self, ctx, coord):
                lb, lg, db, dg = L.bbox_loss(coord.detach().float().contiguous(), tgt, keep)
                if ctx is not None:
                    ctx.g = (db, dg)
                return lb.view(()), lg.view(())

            def bwd(self, ctx, g_b, g_g):
                db, dg = ctx.g
                return (L.axpby_scalars(db, None if g_b is None else _up(g_b), dg, None if g_g is None else _up(g_g)),)
        return self._call(Impl(), output_coord)

    # ------------------------------------------------------------------ MIM
    def get_codebook_indices(self, image):
        """VQKD.get_codebook_indices (model_vqkd.py:173-175): ids int64 [B, num_patches] (no gradient)."""
        self._prep()
        fp, vq = self.flat, self._vq
        B = image.shape[0]
        # model_vqkd.py:125-131 `if data.max() <= 1: data *= 255`: same data-dependent rule, decided on the device (the
        # reference's host branch costs a device->host sync per step)
        pre_mul = torch.where(image.max() <= 1.0, 255.0, 1.0).to(torch.float32).reshape(1)
        y32, _, _ = vq.forward(image, train=False, save=False, pre_mul=pre_mul, pool=False)
        N, D = vq.N, vq.D
        # encode_task_layer (Linear - Tanh - Linear) runs in fp32 in the reference even under autocast
        # (model_vqkd.py:154-155): fp32-grade products on the bf16 tensor cores through the three-term operand split
        w0, w2 = self._task_layer_operands()
        t = L.gemm(L.split_bf16x3(y32.reshape(B * N, D), 0), w0, bias=fp.view32("vqkd.encode_task_layer.0.bias"),
                   out_dtype=torch.float32)
        z = L.gemm(L.split_bf16x3(t, 0, act=1), w2, bias=fp.view32("vqkd.encode_task_layer.2.bias"), out_dtype=torch.float32)
        prow = (torch.arange(B, device=image.device).view(B, 1) * N + 1 + torch.arange(N - 1, device=image.device)).reshape(-1)
        zp = L.gather_rows(z, prow.contiguous())
        ids = L.vq_argmin(zp, fp.view32("vqkd.quantize.embedding.weight"))
        return ids.view(B, N - 1)

    def _task_layer_operands(self):
        """Split-precision copies of the two encode_task_layer weights, rebuilt when the master buffer changes."""
        fp = self.flat
        ver = fp.P._version
        if getattr(self, "_task_ops", None) is None or self._task_ops[0] != ver:
            w0 = L.split_bf16x3(fp.view32("vqkd.encode_task_layer.0.weight").contiguous(), 1)
            w2 = L.split_bf16x3(fp.view32("vqkd.encode_task_layer.2.weight").contiguous(), 1)
            self._task_ops = (ver, w0, w2)
        return self._task_ops[1], self._task_ops[2]

    def get_mim_loss(self, image_embeds_masked, targets, mask_tokens):
        """xfm.py:624-635: CE against VQ-KD ids of the raw image (targets = image), or MSE against the detached
        unmasked embeddings (targets = image_embeds)."""
        self._prep()
        model = self
        B, N, D = image_embeds_masked.shape
        if self.use_vision_tokenizer:
            with torch.no_grad():
                ids = self.get_codebook_indices(targets)
            flat_idx = getattr(mask_tokens, "_xfm_rows", None)
            if flat_idx is None:
                flat_idx = torch.nonzero(mask_tokens.reshape(-1)).reshape(-1)
            flat_idx = flat_idx.to(image_embeds_masked.device)
            rows = (flat_idx // (N - 1)) * N + 1 + flat_idx % (N - 1)
            labels = ids.reshape(-1)[flat_idx].contiguous()
            e16 = _twin(image_embeds_masked)

            class Impl:
                def fwd(self, ctx, emb):
                    x = L.gather_rows(e16.reshape(B * N, D), rows)
                    loss, st = model._mim_head.loss(x, labels)
                    if ctx is not None:
                        ctx.st = st
                    return loss.view(())

                def bwd(self, ctx, g):
                    dx = model._mim_head.backward(ctx.st, _up(g))
                    d = torch.zeros((B * N, D), dtype=torch.float32, device=dx.device)
                    L.scatter_add_rows_(d, rows, dx)
                    return (d.view(B, N, D),)
            return self._call(Impl(), image_embeds_masked)

        tgt = targets.detach().to(torch.float32).contiguous()
        mu8 = mask_tokens.to(torch.uint8).contiguous()

        class ImplMSE:
            def fwd(self, ctx, emb):
                loss, dx = L.mim_mse(emb.detach().to(torch.float32).contiguous(), tgt, mu8, with_cls=not model.mim_cls_only)
                if ctx is not None:
                    ctx.dx = dx
                return loss.view(())

            def bwd(self, ctx, g):
                return (L.scale_by_scalar(ctx.dx, _up(g)),)
        return self._call(ImplMSE(), image_embeds_masked)
