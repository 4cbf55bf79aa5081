"""Rotated RPN inference (detectron2 RRPN, which the reference's RotatedRPN inherits unchanged:
glass/modeling/proposal_generator/rotated_rpn.py:17; hyper-parameters configs/glass_pretrain.yaml:55-74).

ProposalGenerator.forward(images, features) -> per-image (proposal_boxes [<=100,5], objectness_logits).
StandardRPNHead's 3x3 conv + ReLU runs on the tcgen05 conv kernel per level (shared weights); the
objectness and anchor-delta 1x1 convs are fused into ONE GEMM with 72 (padded 80) outputs; anchors,
apply_deltas, per-level top-k, clip, batched rotated NMS and the final top-k all stay on device.
"""
import math
from typing import Dict, List, Sequence, Tuple

import os

import torch

from .. import ops, packing
from ..ops import Act
from ..structures import ImageList, Instances, RotatedBoxes
from .backbone import Workspace


def rotated_cell_anchors(size: float, ratios: Sequence[float], angles: Sequence[float]) -> List[Tuple[float, float, float]]:
    """d2 RotatedAnchorGenerator.generate_cell_anchors (order ratio -> angle), fp32-rounded like torch.tensor()."""
    out = []
    area = size ** 2.0
    for ar in ratios:
        w = math.sqrt(area / ar)
        h = ar * w
        for a in angles:
            out.append((w, h, float(a)))
    return out


class B200RotatedRPN:
    in_features = ("p2", "p3", "p4", "p5", "p6")

    def __init__(self, state_dict: Dict[str, torch.Tensor], prefix: str = "proposal_generator.rpn_head.",
                 anchor_sizes=((16,), (32,), (64,), (128,), (256,)), anchor_ratios=(0.2, 0.5, 1.0),
                 anchor_angles=(-90, -45, 0, 45), strides=(4, 8, 16, 32, 64), bbox_reg_weights=(1.0, 1.0, 1.0, 1.0, 2.0),
                 pre_nms_topk: int = 1000, post_nms_topk: int = 100, nms_thresh: float = 0.7, device="cuda",
                 mode: int = ops.MODE_SPLIT):
        sd = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        self.mode, self.device = mode, device
        self.A = len(anchor_ratios) * len(anchor_angles)
        self.cell_anchors = [rotated_cell_anchors(float(s[0]), anchor_ratios, anchor_angles) for s in anchor_sizes]
        self.strides, self.weights = tuple(strides), tuple(bbox_reg_weights)
        self.pre_nms_topk, self.post_nms_topk, self.nms_thresh = pre_nms_topk, post_nms_topk, nms_thresh
        self.conv = packing.pack_conv(sd["conv.weight"], None, sd["conv.bias"], (1, 1), (1, 1), device=device)
        w = torch.cat((sd["objectness_logits.weight"], sd["anchor_deltas.weight"]), 0)   # [A + 5A, 256, 1, 1]
        b = torch.cat((sd["objectness_logits.bias"], sd["anchor_deltas.bias"]), 0)
        self.pred = packing.pack_conv(w, None, b, (1, 1), (0, 0), n_align=16, device=device)
        # two layers on top of the pyramid: 4 k-blocks per accumulation chunk (logits within 6e-7 of the oracle's)
        self.conv.kb_per_chunk = self.pred.kb_per_chunk = int(os.environ.get("GLASS_KB_HEADS", 4))
        self.ws = Workspace(device)
        self._streams = None

    def head(self, features: Dict[str, Act]) -> List[torch.Tensor]:
        """-> per level fp32 [n, h, w, ld] with columns [0,A) objectness, [A,6A) deltas (tap T3)."""
        preds = []
        for lvl, name in enumerate(self.in_features):
            x = features[name]
            t = self.ws.act(f"rpn.t{lvl}", x.n, 256, x.h, x.w)
            ops.conv2d(x, self.conv, relu=True, out=t, mode=self.mode)
            pred = self.ws.raw(f"rpn.pred{lvl}", (x.n, x.h, x.w, self.pred.n_p), torch.float32)
            ops.conv_gemm(t.hi, t.lo, t.rows, t.cp, [0], self.pred, (t.n, t.hp, t.wp, t.border), out_f32=pred,
                          ld_f32=self.pred.n_p, out_geom=(x.h, x.w, 0), mode=self.mode)
            preds.append(pred)
        return preds

    def topk_decode(self, preds: List[torch.Tensor]):
        """-> (boxes [n, L*K, 5], scores [n, L*K]) per-level top-k decoded proposals (tap T4)."""
        n = preds[0].shape[0]
        L, K = len(preds), self.pre_nms_topk
        boxes = self.ws.raw("rpn.topk_boxes", (n, L * K, 5), torch.float32)
        scores = self.ws.raw("rpn.topk_scores", (n, L * K), torch.float32)
        # The select kernel runs ONE CTA per image, so a level occupies n of the 148 SMs: the five levels are
        # independent (disjoint output slices, own workspaces) and are forked onto side streams so they run side by
        # side, joined before the NMS.  All buffers are persistent workspace, so no allocator/stream hazards arise.
        cur = torch.cuda.current_stream()
        if self._streams is None:
            self._streams = [torch.cuda.Stream() for _ in range(L - 1)]
        for lvl, pred in enumerate(preds):
            need = ops._lib.load().glass_rpn_topk_workspace_bytes(n, pred.shape[1], pred.shape[2], self.A)
            wsb = self.ws.raw(f"rpn.topk_ws{lvl}", (need,), torch.uint8)
            if lvl == 0:
                ops.rpn_topk_decode(pred, self.A, self.strides[lvl], self.cell_anchors[lvl], self.weights, K, lvl, L,
                                    boxes, scores, workspace=wsb)
            else:
                side = self._streams[lvl - 1]
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    ops.rpn_topk_decode(pred, self.A, self.strides[lvl], self.cell_anchors[lvl], self.weights, K, lvl, L,
                                        boxes, scores, workspace=wsb)
        for side in self._streams[: L - 1]:
            cur.wait_stream(side)
        return boxes, scores

    def select(self, boxes: torch.Tensor, scores: torch.Tensor, img_hw: torch.Tensor):
        """clip -> nonempty -> batched rotated NMS (per level) -> first post_nms_topk (tap T5)."""
        n, m, _ = boxes.shape
        L = ops._lib.load()
        wsb = self.ws.raw("rpn.nms_ws", (L.glass_nms_workspace_bytes(n, m),), torch.uint8)
        out = (self.ws.raw("rpn.out_boxes", (n, self.post_nms_topk, 5), torch.float32),
               self.ws.raw("rpn.out_scores", (n, self.post_nms_topk), torch.float32),
               self.ws.raw("rpn.out_index", (n, self.post_nms_topk), torch.int32),
               self.ws.raw("rpn.out_count", (n,), torch.int32))
        return ops.nms_rotated(boxes, scores, self.nms_thresh, self.post_nms_topk, group_size=self.pre_nms_topk,
                               img_hw=img_hw, clip=True, filter_empty=True, workspace=wsb, out=out)

    def forward_device(self, features: Dict[str, Act], img_hw: torch.Tensor):
        """img_hw: fp32 [n,2] device tensor of the (unpadded) image sizes.
        Returns (proposal_boxes [n,100,5], objectness_logits [n,100], index, count [n])."""
        preds = self.head(features)
        boxes, scores = self.topk_decode(preds)
        return self.select(boxes, scores, img_hw)

    @torch.no_grad()
    def forward(self, images: ImageList, features: Dict[str, Act], gt_instances=None):
        """The detectron2 ProposalGenerator surface RotatedRPN inherits (rotated_rpn.py:17; called at
        glass_rcnn.py:88 as ``proposals, _ = self.proposal_generator(images, features, None)``):
        -> (list[Instances{proposal_boxes: RotatedBoxes, objectness_logits}] sorted by score, {} losses)."""
        assert gt_instances is None, "inference only: the training losses are out of scope (DESIGN.md section 7)"
        img_hw = torch.tensor(images.image_sizes, dtype=torch.float32, device=self.device)
        boxes, scores, _, count = self.forward_device(features, img_hw)
        out = []
        for i, c in enumerate(count.cpu().tolist()):
            out.append(Instances(images.image_sizes[i], proposal_boxes=RotatedBoxes(boxes[i, :c].clone()),
                                 objectness_logits=scores[i, :c].clone()))
        return out, {}

    __call__ = forward
