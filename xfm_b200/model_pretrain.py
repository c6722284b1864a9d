"""Pre-training model: the orchestration of models/model_pretrain.py:13-113 (class XFM) on top of xfm_b200.XFMBase.

Same constructor, forward signature, loss dictionary and loss weighting as the reference class; the reference's own
models/model_pretrain.py also runs unchanged against xfm_b200.XFMBase — this mirror exists so
the package is usable where the reference tree is absent (the GPU box, bench.py, tests).
The region / bbox branch (ret_bbox_loss, ret_bbox_giou: idx_to_group_img gather, region-weighted pooling, L1 + GIoU box
losses) follows model_pretrain.py:39-41,81-86.
"""
import os

import torch

from .xfm import XFMBase


class XFM(XFMBase):
    def __init__(self, config, load_vision_params=False, load_text_params=False, **kw):
        super().__init__(config, load_vision_params=load_vision_params, load_text_params=load_text_params,
                         use_contrastive_loss=True, use_matching_loss=True, use_mlm_loss=True,
                         use_bbox_loss=config.get("use_bbox", True), config_text=None, **kw)
        self.weights_map = {"region": config.get("wregion", 1.0), "web": config.get("wweb", 1.0),
                            "imagenet": config.get("wimagenet", 1.0), "image": config.get("wimage", 1.0),
                            "aux": config.get("waux", 1.0)}
        self.do_image_mask = config.get("do_image_mask", True)
        self.use_mm_mim_loss = config.get("use_mm_mim_loss", True)
        self.fuse_itm_mlm = config.get("fuse_itm_mlm", True)  # xfm_b200 only: ITM + MLM in one fusion-encoder pass
        # xfm_b200 only: clean + masked images in one 2B-sample vision pass (XFM_TWIN_VISION=0: A/B measurements)
        self.twin_vision_pass = bool(config.get("twin_vision_pass", os.environ.get("XFM_TWIN_VISION", "1") != "0"))
        self.min_temp = config.get("min_temp", 0.001)
        self.max_temp = config.get("max_temp", 0.5)

    def forward_multimodal(self, image, text_ids, text_atts, text_ids_masked=None, masked_pos=None, masked_ids=None,
                           text_ids_2=None, text_atts_2=None, text_ids_masked_2=None, masked_pos_2=None, masked_ids_2=None,
                           image_atts=None, idx_to_group_img=None, target_bbox=None, is_image=None,
                           ret_mim_loss=False, ret_bbox_loss=False, ret_match_loss=True, ret_mlm_loss=True,
                           ret_bbox_giou=False, ret_itc_loss=True, data_source=None):
        if self.learnable_temp:
            self.clamp_temp(self.min_temp, self.max_temp)
        wmap = self.weights_map
        if ret_bbox_loss:   # region data: fewer images than samples, per-sample region masks (model_pretrain.py:39-41)
            image_embeds, image_atts, image_embeds_fullatts = \
                self.get_vision_embeds(image, image_atts=image_atts, idx_to_group_img=idx_to_group_img)
        else:
            # the masked copy of the same images is asked for below (MIM): both copies run as one 2B-sample encoder pass
            image_embeds, image_atts = self.get_vision_embeds(
                image, _also_masked=bool(ret_mim_loss and self.do_image_mask and self.twin_vision_pass))
        zero = torch.tensor(0.0)
        loss_itc = loss_itm = loss_mlm = loss_mim = loss_bbox = loss_giou = zero
        if data_source != "imagenet":
            both = ret_match_loss and ret_mlm_loss and self.fuse_itm_mlm and text_ids_masked is not None
            text_embeds = self.get_text_embeds(text_ids, text_atts, _also_masked=text_ids_masked if both else None)
            image_feat, text_feat = self.get_features(image_embeds, text_embeds)
            if ret_itc_loss:
                loss_itc = self.get_contrastive_loss(image_feat, text_feat)
                if data_source in wmap:
                    loss_itc = loss_itc * wmap[data_source]
            if ret_match_loss and ret_mlm_loss and self.fuse_itm_mlm:
                # same two losses as the branches below, from one 4B-sample pass of the fusion encoder
                loss_itm, loss_mlm = self.get_matching_and_fuse_mlm_loss(
                    image_embeds, image_atts, image_feat, text_ids, text_atts, text_feat, text_ids_masked, masked_pos,
                    masked_ids, text_embeds=text_embeds)
                if data_source in wmap:
                    loss_itm, loss_mlm = loss_itm * wmap[data_source], loss_mlm * wmap[data_source]
                ret_match_loss = ret_mlm_loss = False
            if ret_match_loss:
                loss_itm = self.get_matching_loss(image_embeds, image_atts, image_feat, text_ids, text_atts, text_feat,
                                                  text_embeds=text_embeds)
                if data_source in wmap:
                    loss_itm = loss_itm * wmap[data_source]
            if ret_mlm_loss:
                loss_mlm = self.get_fuse_mlm_loss(text_ids_masked, text_atts, image_embeds, image_atts, masked_pos, masked_ids)
                if data_source in wmap:
                    loss_mlm = loss_mlm * wmap[data_source]
        if ret_mim_loss and not ret_bbox_loss:
            image_embeds_masked, image_atts, ids_mask = self.get_vision_embeds(image, do_mask=self.do_image_mask)
            if data_source == "imagenet" or self.use_mm_mim_loss:
                target = image if self.use_vision_tokenizer else image_embeds
                loss_mim = self.get_mim_loss(image_embeds_masked, target, ids_mask)
                if data_source in wmap:
                    loss_mim = loss_mim * wmap[data_source]
        if ret_bbox_giou:   # model_pretrain.py:81-86
            output_coord = self.predict_bbox(image_embeds_fullatts, text_ids, text_atts, text_embeds)
            loss_bbox, loss_giou = self.get_bbox_loss(output_coord, target_bbox, is_image=is_image)
        return {"loss_itc": loss_itc, "loss_itm": loss_itm, "loss_mlm": loss_mlm, "loss_mim": loss_mim,
                "loss_bbox": loss_bbox, "loss_giou": loss_giou}

    def forward_text(self, text_ids=None, text_atts=None, text_ids_masked=None, masked_pos=None, masked_ids=None):
        return {"loss_mlm": self.get_mlm_loss(text_ids_masked, text_atts, None, None, masked_pos, masked_ids)}

    def forward(self, image=None, text_ids=None, text_atts=None, text_ids_masked=None, masked_pos=None, masked_ids=None,
                **kw):
        if image is None:
            return self.forward_text(text_ids, text_atts, text_ids_masked, masked_pos, masked_ids)
        return self.forward_multimodal(image, text_ids, text_atts, text_ids_masked, masked_pos, masked_ids, **kw)
