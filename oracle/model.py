"""oracle/model.py -- TEST INFRASTRUCTURE ONLY.

End-to-end CPU restatement of GLASS inference for ``configs/glass_pretrain.yaml``
(SURVEY.md 3.1 / section 8a rows a1-a17), with every parity tap of SURVEY.md B.3
recorded, plus the seeded weight factory (d2-style ``state_dict`` names, A.10) with
BatchNorm calibration.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import d2_ops, nets


@dataclass
class HotPathConfig:
    """Hot-path hyper-parameters (sources: configs/glass_pretrain.yaml,
    glass/config.py, detectron2 defaults)."""
    pixel_mean: Sequence[float] = (103.530, 116.280, 123.675)
    pixel_std: Sequence[float] = (1.0, 1.0, 1.0)
    size_divisibility: int = 32
    anchor_sizes: Sequence[Sequence[float]] = ((16,), (32,), (64,), (128,), (256,))
    anchor_ratios: Sequence[Sequence[float]] = ((0.2, 0.5, 1.0),)
    anchor_angles: Sequence[Sequence[float]] = ((-90, -45, 0, 45),)
    strides: Sequence[int] = (4, 8, 16, 32, 64)
    rpn_bbox_reg_weights: Sequence[float] = (1.0, 1.0, 1.0, 1.0, 2.0)
    rpn_pre_nms_topk: int = 1000
    rpn_post_nms_topk: int = 100
    rpn_nms_thresh: float = 0.7
    box_pooler_resolution: int = 7
    box_pooler_sampling_ratio: int = 2
    box_reg_weights: Sequence[float] = (10.0, 10.0, 5.0, 5.0, 10.0)
    score_thresh_test: float = 0.05
    nms_thresh_test: float = 0.35
    detections_per_image: int = 100
    recog_pool_h: int = 8
    recog_pool_w: int = 32
    recog_sampling_ratio: int = 0
    num_text_classes: int = 97
    max_word_len: int = 26
    max_detections_override: Optional[int] = None  # test-only: cap K to bound CPU time


def _num_anchors(cfg: HotPathConfig) -> int:
    return len(cfg.anchor_sizes[0]) * len(cfg.anchor_ratios[0]) * len(cfg.anchor_angles[0])


class GlassOracle(nn.Module):
    def __init__(self, cfg: Optional[HotPathConfig] = None):
        super().__init__()
        self.cfg = cfg or HotPathConfig()
        self.backbone = nets.ResNetFPN()
        self.proposal_generator = nn.Module()
        self.proposal_generator.rpn_head = nets.RPNHead(_num_anchors(self.cfg), 5)
        rh = nn.Module()
        rh.box_head = nets.FastRCNNConvFCHead(256 * self.cfg.box_pooler_resolution ** 2, 2048)
        rh.box_predictor = nets.RotatedFastRCNNOutputLayers(2048, 1)
        rh.recognizer_feature_fusion = nets.P2P3Fusion(256)
        rh.hybrid_net = nets.ResNetFeatureExtractor(3, 256)
        rh.fusion_net = nets.MultiAspectGCAttention(512, 0.5, 8, 256)
        rh.recognizer_head = nets.RecognizerRCNNHeadV3(256, self.cfg.num_text_classes, self.cfg.max_word_len)
        self.roi_heads = rh
        self.eval()

    # ---------------------------------------------------------------- a1 preprocess
    def preprocess_image(self, images: List[torch.Tensor]):
        """d2 GeneralizedRCNN.preprocess_image + ImageList.from_tensors (A.1);
        called at glass/modeling/meta_arch/glass_rcnn.py:82."""
        mean = torch.tensor(self.cfg.pixel_mean).view(3, 1, 1)
        std = torch.tensor(self.cfg.pixel_std).view(3, 1, 1)
        imgs = [(x.float() - mean) / std for x in images]
        sizes = [(int(im.shape[-2]), int(im.shape[-1])) for im in imgs]
        d = self.cfg.size_divisibility
        mh = (max(s[0] for s in sizes) + d - 1) // d * d
        mw = (max(s[1] for s in sizes) + d - 1) // d * d
        batched = torch.zeros((len(imgs), 3, mh, mw))
        for i, im in enumerate(imgs):
            batched[i, :, : im.shape[-2], : im.shape[-1]] = im
        return batched, sizes

    # ---------------------------------------------------------------- a4 RPN
    def rpn(self, feats: Dict[str, torch.Tensor], image_size: Tuple[int, int], taps=None):
        cfg = self.cfg
        fl = [feats[k] for k in ["p2", "p3", "p4", "p5", "p6"]]
        anchors = d2_ops.rotated_grid_anchors([f.shape[-2:] for f in fl], cfg.strides, cfg.anchor_sizes,
                                              cfg.anchor_ratios, cfg.anchor_angles)
        logits, deltas = self.proposal_generator.rpn_head(fl)
        logits = [s.permute(0, 2, 3, 1).flatten(1) for s in logits]
        deltas = [x.view(x.shape[0], -1, 5, x.shape[-2], x.shape[-1]).permute(0, 3, 4, 1, 2).flatten(1, -2)
                  for x in deltas]
        if taps is not None:
            taps["rpn_logits"] = logits
            taps["rpn_deltas"] = deltas
        proposals = []
        for a, dl in zip(anchors, deltas):
            n = dl.shape[0]
            dl = dl.reshape(-1, 5)
            a = a.unsqueeze(0).expand(n, -1, -1).reshape(-1, 5)
            proposals.append(d2_ops.apply_deltas_rotated(dl, a, cfg.rpn_bbox_reg_weights).view(n, -1, 5))
        res = d2_ops.find_top_rrpn_proposals(proposals, logits, [image_size], cfg.rpn_nms_thresh,
                                             cfg.rpn_pre_nms_topk, cfg.rpn_post_nms_topk, 0.0, taps=taps)
        boxes, scores = res[0]
        if taps is not None:
            taps["proposal_boxes"] = boxes
            taps["objectness_logits"] = scores
        return boxes, scores

    # ---------------------------------------------------------------- a5-a8 box branch
    def box_branch(self, feats, proposal_boxes, image_size, taps=None):
        """recognizers_hybrid_head.py:291-339 + rotated_fast_rcnn.py:88-148,344-373."""
        cfg = self.cfg
        fl = [feats[k] for k in ["p2", "p3", "p4", "p5", "p6"]]
        scales = [1.0 / s for s in cfg.strides]
        pooled = d2_ops.roi_pooler(fl, [proposal_boxes], cfg.box_pooler_resolution, scales,
                                   cfg.box_pooler_sampling_ratio)
        x = self.roi_heads.box_head(pooled)
        scores, deltas, orient = self.roi_heads.box_predictor(x)
        if taps is not None:
            taps["box_pooled"] = pooled
            taps["box_head_out"] = x
            taps["cls_logits"] = scores
            taps["box_deltas"] = deltas
            taps["orient_logits"] = orient
        return self.box_inference(scores, deltas, orient, proposal_boxes, image_size, taps)

    def box_inference(self, scores, deltas, orient, proposal_boxes, image_size, taps=None):
        cfg = self.cfg
        boxes = d2_ops.apply_deltas_rotated(deltas, proposal_boxes, cfg.box_reg_weights)
        probs = F.softmax(scores, dim=-1)
        oprob = F.softmax(orient, dim=-1)
        omax = oprob.max(dim=1)
        orientations = torch.stack((omax[1], omax[0]), 1)  # (index, prob) -> float tensor
        valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(probs).all(dim=1)
        if not valid.all():
            boxes, probs, orientations = boxes[valid], probs[valid], orientations[valid]
        probs = probs[:, :-1]
        boxes = boxes.reshape(-1, 5).clone()
        d2_ops.clip_rotated_(boxes, image_size)
        boxes = boxes.view(-1, 1, 5)
        filter_mask = probs > cfg.score_thresh_test
        filter_inds = torch.nonzero(filter_mask)
        boxes = boxes[filter_inds[:, 0], 0]
        sc = probs[filter_mask]
        orientations = orientations[filter_inds[:, 0]]
        keep = d2_ops.batched_nms_rotated(boxes, sc, filter_inds[:, 1], cfg.nms_thresh_test)
        keep = keep[: cfg.detections_per_image]
        if cfg.max_detections_override is not None:
            keep = keep[: cfg.max_detections_override]
        det = {"pred_boxes": boxes[keep], "scores": sc[keep], "pred_classes": filter_inds[keep][:, 1],
               "orientations": orientations[keep], "kept_proposal_idx": filter_inds[keep][:, 0]}
        if taps is not None:
            taps["det_boxes"] = det["pred_boxes"]
            taps["det_scores"] = det["scores"]
            taps["det_orientations"] = det["orientations"]
        return det

    # ---------------------------------------------------------------- a9-a16 recognizer
    def recognizer_branch(self, images_tensor, feats, det_boxes, taps=None):
        """recognizers_hybrid_head.py:513-569."""
        cfg = self.cfg
        rh = self.roi_heads
        g = rh.recognizer_feature_fusion(feats["p2"], feats["p3"])
        G = d2_ops.roi_pooler([g], [det_boxes], (cfg.recog_pool_h, cfg.recog_pool_w), [1.0 / cfg.strides[0]],
                              cfg.recog_sampling_ratio)
        L0 = d2_ops.roi_pooler([images_tensor], [det_boxes], (cfg.recog_pool_h * 16, cfg.recog_pool_w * 4), [1.0],
                               cfg.box_pooler_sampling_ratio)
        if taps is not None:
            taps["p2p3"] = g
            taps["global_feats"] = G
            taps["local_crops"] = L0
        if L0.shape[0] > 0:
            L = rh.hybrid_net(L0)
            Fcat = torch.cat((L, G), 1)
        else:
            L = None
            Fcat = torch.cat((G, G), 1)
        Fu = rh.fusion_net(Fcat)
        if taps is not None:
            taps["local_feats"] = L
            taps["fusion_out"] = Fu
        if Fu.shape[0] == 0:
            return torch.zeros((0, cfg.max_word_len, cfg.num_text_classes))
        return rh.recognizer_head(Fu, taps=taps)

    # ---------------------------------------------------------------- a17 postprocess
    @staticmethod
    def postprocess(det: Dict[str, torch.Tensor], image_size, out_h, out_w):
        """d2 detector_postprocess / post_processor_academic.py:118-178 (no masks)."""
        sx, sy = out_w / image_size[1], out_h / image_size[0]
        boxes = det["pred_boxes"].clone()
        d2_ops.scale_rotated_(boxes, sx, sy)
        d2_ops.clip_rotated_(boxes, (out_h, out_w))
        keep = d2_ops.nonempty_rotated(boxes)
        out = {k: v[keep] for k, v in det.items() if isinstance(v, torch.Tensor) and v.shape[:1] == keep.shape}
        out["pred_boxes"] = boxes[keep]
        return out

    # ---------------------------------------------------------------- full path
    @torch.no_grad()
    def inference(self, batched_inputs: List[dict], taps: Optional[dict] = None, do_postprocess=True):
        """One image per forward, like the reference."""
        results = []
        for inp in batched_inputs:
            t = {} if taps is not None else None
            images, sizes = self.preprocess_image([inp["image"]])
            feats = self.backbone(images)
            if t is not None:
                t["images"] = images
                for k in ["res2", "res3", "res4", "res5", "p2", "p3", "p4", "p5", "p6"]:
                    t[k] = feats[k]
            pboxes, _ = self.rpn(feats, sizes[0], t)
            det = self.box_branch(feats, pboxes, sizes[0], t)
            det["pred_text_prob"] = self.recognizer_branch(images, feats, det["pred_boxes"], t)
            if t is not None:
                t["pred_text_prob"] = det["pred_text_prob"]
            if do_postprocess:
                det = self.postprocess(det, sizes[0], inp.get("height", sizes[0][0]), inp.get("width", sizes[0][1]))
            results.append({"instances": det})
            if taps is not None:
                taps.setdefault("per_image", []).append(t)
        return results


# ======================================================================================
# Seeded weight factory (SURVEY.md 8c "oracle hygiene", section 0 fact 7)
# ======================================================================================
def _init_weights(model: GlassOracle, seed: int):
    g = torch.Generator().manual_seed(seed)
    bb = model.backbone
    for m in bb.bottom_up.modules():
        if isinstance(m, nn.Conv2d):
            nets.c2_msra_fill(m, g)
    for k in [2, 3, 4, 5]:
        nets.c2_xavier_fill(getattr(bb, f"fpn_lateral{k}"), g)
        nets.c2_xavier_fill(getattr(bb, f"fpn_output{k}"), g)
    rpn = model.proposal_generator.rpn_head
    for m in [rpn.conv, rpn.objectness_logits, rpn.anchor_deltas]:
        nets.normal_fill(m, 0.01, g)
    rh = model.roi_heads
    nets.c2_xavier_fill(rh.box_head.fc1, g)
    nets.c2_xavier_fill(rh.box_head.fc2, g)
    nets.normal_fill(rh.box_predictor.cls_score, 0.01, g)
    nets.normal_fill(rh.box_predictor.bbox_pred, 0.001, g)
    nets.normal_fill(rh.box_predictor.orientation_pred, 0.01, g)
    nets.c2_msra_fill(rh.recognizer_feature_fusion.conv1, g)
    nets.c2_msra_fill(rh.recognizer_feature_fusion.conv2, g)
    for m in list(rh.hybrid_net.modules()) + list(rh.fusion_net.modules()):
        if isinstance(m, nn.Conv2d):
            nets.torch_default_fill(m, g)
    nets.c2_msra_fill(rh.recognizer_head.backbone.conv1, g)
    nets.c2_msra_fill(rh.recognizer_head.backbone.conv2, g)
    for bl in rh.recognizer_head.encoder.bilsm_stack:
        nets.normal_fill(bl.linear, 0.01, g)
        for p in bl.rnn.parameters():
            with torch.no_grad():
                if p.dim() >= 2:
                    nn.init.orthogonal_(p, generator=g)
                else:
                    p.normal_(0, 1, generator=g)
    dec = rh.recognizer_head.decoder.recognizer.decoder
    for m in [dec.attention_unit.sEmbed, dec.attention_unit.xEmbed, dec.attention_unit.wEmbed, dec.fc]:
        nets.torch_default_fill(m, g)
    with torch.no_grad():
        dec.tgt_embedding.weight.normal_(0, 1, generator=g)
        bound = 1.0 / (256 ** 0.5)
        for p in dec.gru.parameters():
            p.uniform_(-bound, bound, generator=g)
        # identity BN/LN hide folding bugs: perturb the affine terms
        for m in model.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.LayerNorm)):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


@torch.no_grad()
def calibrate_batchnorm(model: GlassOracle, images: List[torch.Tensor]):
    """One inference pass with only the BatchNorm modules in train mode, momentum 1:
    running stats <- batch stats of the synthetic input (section 0 fact 7)."""
    bns = [m for m in model.modules() if isinstance(m, nn.BatchNorm2d)]
    for m in bns:
        m.train()
        m.momentum = 1.0
    try:
        model.inference([{"image": im} for im in images], do_postprocess=False)
    finally:
        for m in bns:
            m.eval()
            m.momentum = 0.1


def synthetic_image(seed: int, h: int = 1024, w: int = 1024) -> torch.Tensor:
    """uint8-valued fp32 BGR [3,H,W] image (SURVEY.md 8d cfg 1)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (3, h, w), generator=g).float()


def build_oracle(seed: int = 0, calib_images: Optional[List[torch.Tensor]] = None,
                 cfg: Optional[HotPathConfig] = None) -> GlassOracle:
    model = GlassOracle(cfg)
    _init_weights(model, seed)
    if calib_images is not None:
        calibrate_batchnorm(model, calib_images)
    return model
