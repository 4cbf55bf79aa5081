"""oracle/postprocess.py -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the reference's word post-processor, the step that follows the hot path in every caller
(SURVEY.md 8f #1): ``PostProcessorRotatedBoxes.__call__`` (glass/postprocess/post_processor_rotated_boxes.py:66-86)
followed by ``PostProcessorAcademic.__call__``'s text-score filter (glass/postprocess/post_processor_academic.py:26-34).
Arithmetic is torch fp32 exactly as in the reference; ``cv2.minAreaRect`` (opencv, the third-party routine the
reference calls at post_processor_rotated_boxes.py:261) is called, not restated; rotated IoU / NMS come from the
oracle's C restatement of detectron2's operators (oracle/d2_ops.c).

Pinned by tests/golden/postprocess.pt, produced by the reference's OWN classes
(tools/make_golden_postprocess.py) -- checked by tests/test_oracle_postprocess.py.
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import d2_ops


@dataclass
class PostProcessConfig:
    """cfg.POST_PROCESSING defaults, glass/config.py:176-214."""
    min_box_dim: float = 2.0          # MIN_BOX_DIMENSION
    valid_score: float = 0.15         # VALID_CONFIDENCE
    detect_threshold: float = 0.25    # DETECT_THRESHOLD
    text_threshold: float = 0.25      # TEXT_THRESHOLD
    merge_ioa_thresh: float = 0.3     # MERGE_IOA_THRESH
    pairs_height_ratio_thresh: float = 0.35  # PAIRS_HEIGHT_RATIO_THRESH
    max_angle_diff: float = 15.0      # MAX_ANGLE_DIFF
    minimal_ioa_thresh: float = 0.01  # post_processor_rotated_boxes.py:40
    nms_iou: float = 0.99             # post_processor_rotated_boxes.py:181


def pairwise_ioa_rotated(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """glass/structures/boxes.py:24-49: intersection over the smaller area, recovered from the IoU."""
    iou = d2_ops.box_iou_rotated(b1, b2)
    area1 = b1[:, 2] * b1[:, 3]
    area2 = b2[:, 2] * b2[:, 3]
    a1 = area1.repeat(len(b2), 1).T
    a2 = area2.repeat(len(b1), 1)
    inter = (a1 + a2) * iou / (1 + iou)
    return inter / torch.min(a1, a2)


def boxes_to_polygons(boxes: torch.Tensor) -> torch.Tensor:
    """post_processor_rotated_boxes.py:221-249 (vertex 0 = top-left of the un-rotated box)."""
    n = len(boxes)
    if n == 0:
        return torch.zeros((0, 4, 2), dtype=boxes.dtype)
    cx, cy, w, h, a = boxes.T
    t = (-a / 180) * np.pi
    poly = torch.zeros((n, 4, 2), dtype=boxes.dtype)
    sin_t, cos_t = torch.sin(t), torch.cos(t)
    poly[:, 0, 0] = cx + (h * sin_t - w * cos_t) / 2
    poly[:, 1, 0] = cx + (h * sin_t + w * cos_t) / 2
    poly[:, 2, 0] = cx - (h * sin_t - w * cos_t) / 2
    poly[:, 3, 0] = cx - (h * sin_t + w * cos_t) / 2
    poly[:, 0, 1] = cy - (h * cos_t + w * sin_t) / 2
    poly[:, 1, 1] = cy - (h * cos_t - w * sin_t) / 2
    poly[:, 2, 1] = cy + (h * cos_t + w * sin_t) / 2
    poly[:, 3, 1] = cy + (h * cos_t - w * sin_t) / 2
    return poly


def polygons_to_rotated_boxes(polygons: torch.Tensor, orientations: Optional[torch.Tensor]) -> torch.Tensor:
    """post_processor_rotated_boxes.py:251-286: min-area rectangle of the vertices, re-oriented to the quadrant of
    ``orientations``.  NB the reference passes ``orientations`` in RADIANS (see merge_rotated_boxes) but compares
    them with degrees here; that quirk is part of the behaviour being matched."""
    import cv2
    out = torch.zeros((len(polygons), 5))
    for i, poly in enumerate(polygons.cpu().numpy()):
        center, shape, angle = cv2.minAreaRect(np.array(poly))
        angle = 90 - angle
        diff = (orientations[i] - angle) if orientations is not None else 0.
        diff = (diff + 180) % 360 - 180
        if -45 < diff <= 45:
            width, height = shape[1], shape[0]
        elif 45 < diff <= 135:
            width, height = shape[0], shape[1]
            angle += 90
        elif -135 < diff <= -45:
            width, height = shape[0], shape[1]
            angle -= 90
        else:
            width, height = shape[1], shape[0]
            angle += 180
        angle = (angle + 180) % 360 - 180
        out[i] = torch.tensor([center[0], center[1], width, height, angle])
    return out.to(dtype=polygons.dtype)


def merge_rotated_boxes(b1: torch.Tensor, b2: torch.Tensor, s1: torch.Tensor, s2: torch.Tensor) -> torch.Tensor:
    """post_processor_rotated_boxes.py:186-218 with scores given: orientation of the higher-scored box."""
    p1, p2 = boxes_to_polygons(b1), boxes_to_polygons(b2)
    a1 = b1[:, 4] * np.pi / 180
    a2 = b2[:, 4] * np.pi / 180
    merged_angle = torch.where(s1 >= s2, a1, a2)
    return polygons_to_rotated_boxes(torch.hstack((p1, p2)), merged_angle)


def merge_intersecting_boxes(boxes: torch.Tensor, scores: torch.Tensor, idx: torch.Tensor, cfg: PostProcessConfig,
                             max_iters: int = 1000) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]:
    """post_processor_rotated_boxes.py:107-184.  Returns (boxes, scores, original indices, iterations)."""
    iters = 0
    if len(boxes) == 0:
        return boxes, scores, idx, iters
    boxes = boxes.clone()
    while iters < max_iters:
        ioa = pairwise_ioa_rotated(boxes, boxes)
        pairs = torch.nonzero(ioa.fill_diagonal_(0).triu() >= cfg.minimal_ioa_thresh)
        if len(pairs) == 0:
            break
        h, ang = boxes[:, 3], boxes[:, 4]
        dang = ang[pairs[:, 1]] - ang[pairs[:, 0]]
        dang = torch.abs((dang + 180) % 360 - 180)
        similar_angle = (dang < cfg.max_angle_diff) | (dang > (180 - cfg.max_angle_diff))
        ratio = h[pairs[:, 1]] / h[pairs[:, 0]]
        similar_height = (cfg.pairs_height_ratio_thresh < ratio) & (ratio < (1 / (cfg.pairs_height_ratio_thresh + 1e-6)))
        valid_score = torch.min(scores[pairs[:, 0]], scores[pairs[:, 1]]) >= cfg.valid_score
        ioa_mask = ioa[pairs[:, 0], pairs[:, 1]] >= cfg.merge_ioa_thresh
        ok = valid_score & similar_height & ioa_mask & similar_angle
        if (~ok).all():
            break
        vp = pairs[ok]
        merged = merge_rotated_boxes(boxes[vp[:, 0]], boxes[vp[:, 1]], scores[vp[:, 0]], scores[vp[:, 1]])
        # index_put with repeated indices on CPU: the last pair that names a box wins, first-members first
        for k in range(len(vp)):
            boxes[vp[k, 0]] = merged[k]
        for k in range(len(vp)):
            boxes[vp[k, 1]] = merged[k]
        keep = d2_ops.nms_rotated(boxes, scores, cfg.nms_iou)
        boxes, scores, idx = boxes[keep], scores[keep], idx[keep]
        iters += 1
    return boxes, scores, idx, iters


def post_process(boxes: torch.Tensor, scores: torch.Tensor, text_scores: Optional[torch.Tensor] = None,
                 cfg: PostProcessConfig = PostProcessConfig()):
    """-> (boxes [M,5], original indices [M] int64, polygons [M,4,2], merge iterations)."""
    boxes, scores = boxes.float().clone(), scores.float().clone()
    idx = torch.arange(len(boxes))
    k = torch.min(boxes[:, 2], boxes[:, 3]) >= cfg.min_box_dim if len(boxes) else torch.zeros(0, dtype=torch.bool)
    boxes, scores, idx = boxes[k], scores[k], idx[k]                      # filter_small_boxes :87-92
    k = scores >= cfg.valid_score                                         # :98
    boxes, scores, idx = boxes[k], scores[k], idx[k]
    boxes, scores, idx, iters = merge_intersecting_boxes(boxes, scores, idx, cfg)
    k = scores >= cfg.detect_threshold                                    # :103
    boxes, scores, idx = boxes[k], scores[k], idx[k]
    if text_scores is not None:                                           # post_processor_academic.py:31-32
        k = text_scores.float()[idx] >= cfg.text_threshold
        boxes, scores, idx = boxes[k], scores[k], idx[k]
    return boxes, idx, boxes_to_polygons(boxes), iters


# ----------------------------------------------------------------------------------------------------------------
# a17: GlassRCNN._postprocess (glass/modeling/meta_arch/glass_rcnn.py:103-128) and the detector_postprocess it ends in
# (glass/postprocess/post_processor_academic.py:118-178).  Pinned by tests/golden/meta_postprocess.pt, written by the
# reference's OWN functions (tools/make_golden_meta_postprocess.py) -- tests/test_oracle_meta_postprocess.py.
# ----------------------------------------------------------------------------------------------------------------
def filter_small_boxes_keep(boxes: torch.Tensor, min_box_dim: float) -> torch.Tensor:
    """post_processor_rotated_boxes.py:89-94 -> boolean keep mask."""
    return torch.min(boxes[:, 2], boxes[:, 3]) >= min_box_dim


def resize_boxes_(boxes: torch.Tensor, ratio: float, image_size, axis: str = "both") -> torch.Tensor:
    """post_processor_academic.py:36-63, in place: w += ratio*w, h += ratio*h, then RotatedBoxes.clip."""
    if len(boxes) == 0:
        return boxes
    dx = ratio * boxes[:, 2] if axis in ("both", "horizontal") else 0
    dy = ratio * boxes[:, 3] if axis in ("both", "vertical") else 0
    boxes[:, 2] += dx
    boxes[:, 3] += dy
    d2_ops.clip_rotated_(boxes, image_size)
    return boxes


def detector_postprocess(boxes: torch.Tensor, image_size, out_h: int, out_w: int, rboxes_alias: bool = False):
    """post_processor_academic.py:118-178 on a box tensor: scale, clip, nonempty.  Returns (boxes, keep[, rboxes]).
    ``rboxes_alias``: ``pred_rboxes`` is the SAME object as ``pred_boxes`` (recognizers_hybrid_head.py:596-597), so the
    in-place scale/clip of :158-159 already moved it and :173-175 scale and clip it a second time."""
    sx, sy = out_w / image_size[1], out_h / image_size[0]
    b = boxes.clone()
    d2_ops.scale_rotated_(b, sx, sy)
    d2_ops.clip_rotated_(b, (out_h, out_w))
    keep = d2_ops.nonempty_rotated(b)
    out = b[keep]
    if not rboxes_alias:
        return out, keep
    rb = out.clone()
    d2_ops.scale_rotated_(rb, sx, sy)
    d2_ops.clip_rotated_(rb, (out_h, out_w))
    return out, keep, rb


def glass_rcnn_postprocess(boxes: torch.Tensor, image_size, out_h: int, out_w: int,
                           min_box_dim: Optional[float] = None, inflate_ratio: Optional[float] = None):
    """glass_rcnn.py:103-128 for one image -> (boxes, original indices of the survivors)."""
    idx = torch.arange(len(boxes))
    b = boxes.clone()
    if min_box_dim and len(b):
        k = filter_small_boxes_keep(b, min_box_dim)
        b, idx = b[k], idx[k]
    if inflate_ratio:
        resize_boxes_(b, inflate_ratio, image_size)
    out, keep = detector_postprocess(b, image_size, out_h, out_w)
    return out, idx[keep]
