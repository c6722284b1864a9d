"""NLVR2 fine-tuning model: the orchestration of models/model_nlvr.py:16-44 (class XFMForNLVR) on xfm_b200.XFMBase — two
images per text through the cross-attention fusion encoder, concatenated CLS rows, build_mlp head, cross-entropy.
The head's parameters live in the flat buffers like every other parameter (same `cls_head.N.*` names as the reference's
nn.Sequential), so the flat optimizer updates them and the accelerator reduces / clips / zeroes their gradients."""
import torch
import torch.nn.functional as F

from . import encoders as E
from .xfm import XFMBase


class XFMForNLVR(XFMBase):
    def __init__(self, config, **kw):
        super().__init__(config, load_vision_params=False, load_text_params=False, use_contrastive_loss=False,
                         use_matching_loss=False, use_mlm_loss=False, use_bbox_loss=False, config_text=None, **kw)
        if "load_domain_pretrained" not in config or not config["load_domain_pretrained"]:   # model_nlvr.py:26
            self.init_params = ["cls_head." + n for n in ("0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias")]

    def _extend_params(self, fp, cfg, init, config):
        self._head_names["cls_head"] = 2
        E.add_mlp_head(fp, init, "cls_head", cfg["hidden"] * 2, 2)   # model_nlvr.py:25 build_mlp(text_width * 2, 2)

    def forward(self, image, text_ids, text_atts, targets, train=True):
        image_embeds, image_atts = self.get_vision_embeds(image)
        encoder_embeds = self.get_text_embeds(text_ids, text_atts)
        image0_embeds, image1_embeds = torch.split(image_embeds, targets.size(0))
        cls1 = self.get_cross_embeds(image0_embeds, image_atts[:image0_embeds.size(0)], text_embeds=encoder_embeds,
                                     text_atts=text_atts, is_pretrain=False)[:, 0, :]
        cls2 = self.get_cross_embeds(image1_embeds, image_atts[image0_embeds.size(0):], text_embeds=encoder_embeds,
                                     text_atts=text_atts, is_pretrain=False)[:, 0, :]
        output_cls = torch.cat((cls1, cls2), dim=-1)
        assert output_cls.shape[-1] == self.text_width * 2
        prediction = self.cls_head(output_cls)
        return F.cross_entropy(prediction, targets) if train else prediction
