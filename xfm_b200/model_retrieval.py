"""Retrieval fine-tuning model: the orchestration of models/model_retrieval.py:11-37 (class XFMForRetrieval) on
xfm_b200.XFMBase — ITC with soft labels from `idx` (xfm.py:705-713) + hard-negative ITM with `idx`-masked sampling
(xfm.py:729-734), text gradients flowing into the fusion encoder (is_pretrain=False, xfm.py:674).  The reference's own file
also runs unchanged against xfm_b200.XFMBase; this mirror serves tests / benchmarks where the reference
tree is absent."""
from .xfm import XFMBase, load_pretrained


class XFMForRetrieval(XFMBase):
    def __init__(self, config, **kw):
        super().__init__(config, load_vision_params=False, load_text_params=False, use_contrastive_loss=True,
                         use_matching_loss=True, use_mlm_loss=False, use_bbox_loss=False, **kw)
        self.init_params = []

    def load_pretrained(self, ckpt_rpath, config, is_eval=False):
        state_dict = load_pretrained(self, ckpt_rpath, config, is_eval=is_eval, load_text=True)
        msg = self.load_state_dict(state_dict, strict=False)
        print("load checkpoint from %s" % ckpt_rpath)
        print("missing_keys: ", [p for p in msg.missing_keys if "vision_encoder" not in p])
        print("unexpected_keys: ", msg.unexpected_keys)

    def forward(self, image, text_ids, text_atts, idx=None):
        image_embeds, image_atts = self.get_vision_embeds(image)
        text_embeds = self.get_text_embeds(text_ids, text_atts)
        image_feat, text_feat = self.get_features(image_embeds, text_embeds)
        loss_itc = self.get_contrastive_loss(image_feat, text_feat, idx=idx)
        loss_itm = self.get_matching_loss(image_embeds, image_atts, image_feat, text_ids, text_atts, text_feat, idx=idx,
                                          text_embeds=text_embeds, is_pretrain=False)
        return loss_itc, loss_itm
