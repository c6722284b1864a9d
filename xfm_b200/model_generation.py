"""VQA fine-tuning model (BASELINE config #5): the orchestration of models/model_generation.py:23-202 (class XFMForVQA) on
xfm_b200.XFMBase plus the cross-modal CAUSAL decoder the reference builds from models/xroberta.py:963-1123
(RobertaForCausalLM with cross-attention over the question states in every layer, decoder_fusion_start_at 0).

The decoder runs on the same kernels as the fusion encoder (blocks.roberta_layer_fwd / _bwd): the causal mask
(xroberta.py:771-806) is an additive [H, La, La] term of the self-attention, the question padding mask
(xroberta.py:903-909) an additive key mask of the cross-attention, and every answer indexes the question whose states it
attends to (kv_index) — the reference materialises `question_states` / `question_atts` repeated per answer
(model_generation.py:111-117) and recomputes the cross K/V projection for each copy; here it runs once per question.
The LM head + shifted CrossEntropyLoss(reduction='none') summed per sequence (xroberta.py:1104-1110) is one fused op with a
per-row upstream gradient in its backward (xfm_ce_bwd_rows)."""
import torch
import torch.nn.functional as F

from . import encoders as E
from . import lib as L
from .xfm import XFMBase, _twin, load_pretrained


class XFMForVQA(XFMBase):
    """Generative model; inference ranks a fixed answer list (model_generation.py:24-27)."""

    def __init__(self, config, **kw):
        super().__init__(config, load_vision_params=False, load_text_params=False, use_contrastive_loss=False,
                         use_matching_loss=False, use_mlm_loss=False, use_bbox_loss=False, config_text=None, **kw)
        assert isinstance(config["pad_token_id"], int)
        self.pad_token_id = config["pad_token_id"]
        self.cross_encoder_width, self.dec_encoder_width = self.vision_width, self.text_width
        if self.dec_encoder_width != self.cross_encoder_width:  # model_generation.py:56-60
            self.init_params = [n for n in self._params if n.startswith("text_decoder.") and
                                ("crossattention.self.key" in n or "crossattention.self.value" in n)]
        else:
            self.init_params = []

    # ------------------------------------------------------------------ parameters / runners
    def _extend_params(self, fp, cfg, init, config):
        if config.get("decoder_fusion_start_at", 0) != 0:
            raise NotImplementedError("the decoder is built with cross-attention in every layer (decoder_fusion_start_at 0)")
        self.dec_layers = int(config.get("num_dec_layers", config.get("dec_layers", cfg["fusion_layers"])))
        E.add_roberta(fp, cfg, init, "text_decoder.", self.dec_layers, cross=True, enc_width=cfg["hidden"], heads=("lm_head",))

    def _extend_modules(self, fp, cfg, config):
        d = "text_decoder."
        self._register(d + "lm_head.decoder.weight", self._params[d + "roberta.embeddings.word_embeddings.weight"])
        self._register(d + "lm_head.decoder.bias", self._params[d + "lm_head.bias"])
        self._register_buffer(d + "roberta.embeddings.position_ids", torch.arange(cfg["max_pos"]).expand((1, -1)).clone())
        self._dec = E.RobertaStack(fp, cfg, d, self.dec_layers, cross=True)
        self._dec_head = E.LMHead(fp, cfg, d)
        self._causal = {}

    def _extra_runners(self):
        self._causal = {}
        return (self._dec,)

    def load_pretrained(self, ckpt_rpath, config, is_eval=False):
        """model_generation.py:62-91: the decoder starts from the pre-trained fusion encoder."""
        if is_eval:
            state_dict = load_pretrained(self, ckpt_rpath, config, is_eval=True)
        else:
            state_dict = load_pretrained(self, ckpt_rpath, config, load_text=False)
            for key in list(state_dict.keys()):
                if "fusion_encoder." in key:
                    state_dict[key.replace("fusion_encoder", "text_decoder")] = state_dict[key]
        msg = self.load_state_dict(state_dict, strict=False)
        print("load checkpoint from %s" % ckpt_rpath)
        print("missing_keys: ", [p for p in msg.missing_keys if "vision_encoder" not in p])
        print("unexpected_keys: ", msg.unexpected_keys)

    # ------------------------------------------------------------------ decoder
    def _causal_bias(self, La, device):
        """Additive causal term, f32 [H, La, ld] (ld even): 0 where key <= query, -10000 elsewhere (xroberta.py:775-806)."""
        key = (La, str(device))
        if key not in self._causal:
            H, ld = self.cfg["heads"], (La + 1) // 2 * 2
            ids = torch.arange(ld, device=device)
            m = (ids[None, :] > torch.arange(La, device=device)[:, None]).to(torch.float32) * -10000.0
            self._causal[key] = m.unsqueeze(0).expand(H, La, ld).contiguous()
        return self._causal[key]

    def _decode(self, answer_ids, answer_atts, question_states, question_atts, kv_index, labels=None):
        """text_decoder(...) of model_generation.py:119-127,150-154,180-186 on answers [Na, La] whose sample a attends to
        question_states[kv_index[a]].  question_atts: [Bq, Lq] or None (all ones).  With `labels` returns the per-answer
        next-token losses [Na] (reduction='none', summed per sequence); without, the f32 logits [Na, La, V]."""
        self._prep()
        model = self
        Na, La = answer_ids.shape
        Bq, Lq, D = question_states.shape
        dev = answer_ids.device
        q16 = _twin(question_states)
        kv_index = kv_index.to(torch.int32).contiguous()
        kmask = E.RobertaStack.additive_mask(answer_atts if answer_atts is not None else torch.ones_like(answer_ids))
        enc_kmask = None
        if question_atts is not None:
            enc_kmask = E.RobertaStack.additive_mask(question_atts)[kv_index.long()].contiguous()
        self_bias = self._causal_bias(La, dev) if La > 1 else None
        V = self.cfg["vocab_size"]
        if labels is not None:
            rows = (torch.arange(Na, device=dev).view(Na, 1) * La + torch.arange(La - 1, device=dev).view(1, -1)).reshape(-1)
            tgt = labels[:, 1:].reshape(-1).contiguous()

        class Impl:
            def fwd(self, ctx, q):
                drop = model._drop()
                h, h32, est = model._dec.embed(answer_ids, drop, save=self.save)
                h, h32, st = model._dec.layers_fwd(h, Na, La, kmask, enc=q16.reshape(Bq * Lq, D), Benc=Bq, Lenc=Lq,
                                                   kv_index=kv_index, drop=drop, save=self.save, h32=h32,
                                                   self_bias=self_bias, enc_kmask=enc_kmask)
                if labels is None:
                    return model._dec_head.logits(h)[:, :V].reshape(Na, La, V)
                row_loss, hst = model._dec_head.loss_rows(L.gather_rows(h, rows), tgt)
                if ctx is not None:
                    ctx.est, ctx.st, ctx.hst = est, st, hst
                return row_loss.view(Na, La - 1).sum(1)

            def bwd(self, ctx, g):
                scale = g.to(torch.float32).reshape(Na, 1).expand(Na, La - 1).reshape(-1).contiguous()
                dx = model._dec_head.backward(ctx.hst, None, row_scale=scale)
                dh = torch.zeros((Na * La, D), dtype=torch.float32, device=dx.device)
                L.scatter_add_rows_(dh, rows, dx)
                d_enc = torch.zeros((Bq * Lq, D), dtype=torch.float32, device=dx.device)
                dh = model._dec.layers_bwd(ctx.st, dh, d_enc=d_enc, need_dh=True)
                model._dec.embed_bwd(ctx.est, dh)
                ctx.est = ctx.st = ctx.hst = None
                return (d_enc.view(Bq, Lq, D),)
        if labels is None:
            with torch.no_grad():
                return self._call(Impl(), question_states)
        return self._call(Impl(), question_states)

    # ------------------------------------------------------------------ model_generation.py:93-144
    def forward(self, image, quesiton, answer=None, k=None, weights=None, train=True):
        image_embeds, image_atts = self.get_vision_embeds(image)
        text_embeds = self.get_text_embeds(quesiton.input_ids, quesiton.attention_mask)
        question_output = self.get_cross_embeds(image_embeds, image_atts, text_embeds=text_embeds,
                                                text_atts=quesiton.attention_mask, is_pretrain=False)
        if train:
            # k: number of answers for each question; weights: weight for each answer
            answer_targets = answer.input_ids.masked_fill(answer.input_ids == self.pad_token_id, -100)
            counts = k if torch.is_tensor(k) else torch.as_tensor(k, device=image.device)   # (a device tensor avoids a copy)
            kv_index = torch.repeat_interleave(torch.arange(len(k), device=image.device), counts,
                                               output_size=answer.input_ids.shape[0])
            answer_loss = self._decode(answer.input_ids, answer.attention_mask, question_output, quesiton.attention_mask,
                                       kv_index, labels=answer_targets)
            self.last_answer_loss = answer_loss.detach()
            loss = weights * answer_loss
            return loss.sum() / image.size(0)
        return self.rank_answer(question_output, None, answer.input_ids, answer.attention_mask, k)

    def rank_answer(self, question_states, question_atts, answer_ids, answer_atts, k):
        """model_generation.py:146-202.  question_atts None = all ones (what forward passes at inference, :141)."""
        with torch.no_grad():
            num_ques = question_states.size(0)
            dev = answer_ids.device
            each = torch.arange(num_ques, device=dev)
            start_ids = answer_ids[0, 0].repeat(num_ques, 1)  # bos token
            logits = self._decode(start_ids, None, question_states, question_atts, each)[:, 0, :]
            answer_first_token = answer_ids[:, 1]
            prob_first_token = F.softmax(logits, dim=1).index_select(dim=1, index=answer_first_token)
            topk_probs, topk_ids = prob_first_token.topk(k, dim=1)
            flat = topk_ids.reshape(-1)
            input_ids = answer_ids.index_select(0, flat)
            input_atts = answer_atts.index_select(0, flat)
            targets_ids = input_ids.masked_fill(input_ids == self.pad_token_id, -100)
            # the reference tiles the question states k times (:176-177); here answer a indexes question a // k
            answer_loss = self._decode(input_ids, input_atts, question_states, question_atts,
                                       torch.repeat_interleave(each, k), labels=targets_ids)
            log_probs = torch.cat([topk_probs.view(-1, 1).log(), -answer_loss.view(-1, 1)], dim=1)
            log_probs_sum = log_probs.sum(1).view(num_ques, k)  # chain rule over the answer tokens
            topk_probs = F.softmax(log_probs_sum, dim=-1)
            topk_probs, rerank_id = topk_probs.topk(k, dim=1)
            topk_ids = torch.gather(topk_ids, 1, rerank_id)
            return topk_ids, topk_probs
