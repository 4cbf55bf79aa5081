"""oracle/d2_ops.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the detectron2 v0.6 [d2-recall] operators / helpers that the
reference calls on its inference path.  The three native
operators are in ``d2_ops.c`` (loaded here through ctypes); everything else is plain
torch.  Each function cites the reference call site it serves.
"""
import ctypes
import math
from typing import List, Sequence, Tuple

import torch

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.oracle_single_box_iou_rotated.restype = ctypes.c_float
        _lib.oracle_nms_rotated.restype = ctypes.c_int
    return _lib


def _fp(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


# --------------------------------------------------------------------------------------
# ROIAlignRotated (d2 layers/roi_align_rotated.py + csrc/ROIAlignRotated) -- A.6
# --------------------------------------------------------------------------------------
def roi_align_rotated(input: torch.Tensor, rois: torch.Tensor, output_size: Tuple[int, int],
                      spatial_scale: float, sampling_ratio: int) -> torch.Tensor:
    """input [N,C,H,W] fp32, rois [M,6]=(batch,cx,cy,w,h,deg) -> [M,C,oh,ow].

    Call sites: recognizers_hybrid_head.py:320 (box), :550 (recognizer), :556 (image)."""
    assert input.dtype == torch.float32 and rois.shape[1] == 6
    input = input.contiguous()
    rois = rois.contiguous().float()
    oh, ow = output_size
    m = rois.shape[0]
    out = torch.zeros((m, input.shape[1], oh, ow), dtype=torch.float32)
    if m == 0:
        return out
    lib().oracle_roi_align_rotated_forward(
        _fp(input), _fp(rois), _fp(out), ctypes.c_int(m), ctypes.c_int(input.shape[1]),
        ctypes.c_int(input.shape[2]), ctypes.c_int(input.shape[3]), ctypes.c_int(oh),
        ctypes.c_int(ow), ctypes.c_float(spatial_scale), ctypes.c_int(sampling_ratio))
    return out


# --------------------------------------------------------------------------------------
# rotated IoU / NMS (d2 csrc/box_iou_rotated, csrc/nms_rotated, layers/nms.py) -- A.4
# --------------------------------------------------------------------------------------
def box_iou_rotated(boxes1: torch.Tensor, boxes2: torch.Tensor) -> torch.Tensor:
    b1 = boxes1.contiguous().float()
    b2 = boxes2.contiguous().float()
    out = torch.zeros((b1.shape[0], b2.shape[0]), dtype=torch.float32)
    if out.numel():
        lib().oracle_box_iou_rotated(_fp(b1), ctypes.c_int(b1.shape[0]), _fp(b2),
                                     ctypes.c_int(b2.shape[0]), _fp(out))
    return out


def nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Greedy rotated NMS; returns kept indices in descending-score order (int64)."""
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64)
    b = boxes.contiguous().float()
    # stable sort so ties resolve by index (the device path uses a stable radix sort)
    order = torch.sort(scores.float(), descending=True, stable=True)[1].contiguous()
    keep = torch.empty((n,), dtype=torch.int64)
    k = lib().oracle_nms_rotated(_fp(b), _fp(order), ctypes.c_int(n),
                                 ctypes.c_float(iou_threshold), _fp(keep))
    return keep[:k].clone()


def batched_nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor,
                        iou_threshold: float) -> torch.Tensor:
    """d2 layers/nms.py batched_nms_rotated (rotated_fast_rcnn.py:131; RRPN proposals)."""
    assert boxes.shape[-1] == 5
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    boxes = boxes.float()
    max_coordinate = (torch.max(boxes[:, 0], boxes[:, 1]) + torch.max(boxes[:, 2], boxes[:, 3]) / 2).max()
    min_coordinate = (torch.min(boxes[:, 0], boxes[:, 1]) - torch.max(boxes[:, 2], boxes[:, 3]) / 2).min()
    offsets = idxs.to(boxes) * (max_coordinate - min_coordinate + 1)
    boxes_for_nms = boxes.clone()
    boxes_for_nms[:, :2] += offsets[:, None]
    return nms_rotated(boxes_for_nms, scores, iou_threshold)


# --------------------------------------------------------------------------------------
# RotatedBoxes helpers (d2 structures/rotated_boxes.py) -- A.4
# --------------------------------------------------------------------------------------
def normalize_angles_(t: torch.Tensor) -> torch.Tensor:
    t[:, 4] = (t[:, 4] + 180.0) % 360.0 - 180.0
    return t


def clip_rotated_(t: torch.Tensor, box_size: Tuple[int, int], clip_angle_threshold: float = 1.0) -> torch.Tensor:
    h, w = box_size
    normalize_angles_(t)
    idx = torch.where(torch.abs(t[:, 4]) <= clip_angle_threshold)[0]
    x1 = t[idx, 0] - t[idx, 2] / 2.0
    y1 = t[idx, 1] - t[idx, 3] / 2.0
    x2 = t[idx, 0] + t[idx, 2] / 2.0
    y2 = t[idx, 1] + t[idx, 3] / 2.0
    x1.clamp_(min=0, max=w)
    y1.clamp_(min=0, max=h)
    x2.clamp_(min=0, max=w)
    y2.clamp_(min=0, max=h)
    t[idx, 0] = (x1 + x2) / 2.0
    t[idx, 1] = (y1 + y2) / 2.0
    t[idx, 2] = torch.min(t[idx, 2], x2 - x1)
    t[idx, 3] = torch.min(t[idx, 3], y2 - y1)
    return t


def nonempty_rotated(t: torch.Tensor, threshold: float = 0.0) -> torch.Tensor:
    return (t[:, 2] > threshold) & (t[:, 3] > threshold)


def scale_rotated_(t: torch.Tensor, scale_x: float, scale_y: float) -> torch.Tensor:
    t[:, 0] *= scale_x
    t[:, 1] *= scale_y
    theta = t[:, 4] * math.pi / 180.0
    c = torch.cos(theta)
    s = torch.sin(theta)
    t[:, 2] *= torch.sqrt((scale_x * c) ** 2 + (scale_y * s) ** 2)
    t[:, 3] *= torch.sqrt((scale_x * s) ** 2 + (scale_y * c) ** 2)
    t[:, 4] = torch.atan2(scale_x * s, scale_y * c) * 180 / math.pi
    return t


# --------------------------------------------------------------------------------------
# Box2BoxTransformRotated.apply_deltas (d2 modeling/box_regression.py) -- A.4
#   call sites: RRPN decode (weights 1,1,1,1,2: glass_pretrain.yaml:66);
#   rotated_fast_rcnn.py:335-342 (weights 10,10,5,5,10: glass_pretrain.yaml:98)
# --------------------------------------------------------------------------------------
SCALE_CLAMP = math.log(1000.0 / 16)


def apply_deltas_rotated(deltas: torch.Tensor, boxes: torch.Tensor, weights: Sequence[float]) -> torch.Tensor:
    assert deltas.shape[1] % 5 == 0 and boxes.shape[1] == 5
    boxes = boxes.to(deltas.dtype).unsqueeze(2)
    ctr_x, ctr_y, widths, heights, angles = (boxes[:, i] for i in range(5))
    wx, wy, ww, wh, wa = weights
    dx = deltas[:, 0::5] / wx
    dy = deltas[:, 1::5] / wy
    dw = deltas[:, 2::5] / ww
    dh = deltas[:, 3::5] / wh
    da = deltas[:, 4::5] / wa
    dw = torch.clamp(dw, max=SCALE_CLAMP)
    dh = torch.clamp(dh, max=SCALE_CLAMP)
    pred = torch.zeros_like(deltas)
    pred[:, 0::5] = dx * widths + ctr_x
    pred[:, 1::5] = dy * heights + ctr_y
    pred[:, 2::5] = torch.exp(dw) * widths
    pred[:, 3::5] = torch.exp(dh) * heights
    pred_angle = da * 180.0 / math.pi + angles
    pred_angle = (pred_angle + 180.0) % 360.0 - 180.0
    pred[:, 4::5] = pred_angle
    return pred


# --------------------------------------------------------------------------------------
# RotatedAnchorGenerator (d2 modeling/anchor_generator.py) -- A.4
#   config: glass_pretrain.yaml:55-59
# --------------------------------------------------------------------------------------
def rotated_cell_anchors(sizes: Sequence[float], aspect_ratios: Sequence[float],
                         angles: Sequence[float]) -> torch.Tensor:
    anchors = []
    for size in sizes:
        area = size ** 2.0
        for ar in aspect_ratios:
            w = math.sqrt(area / ar)
            h = ar * w
            anchors.extend([0, 0, w, h, a] for a in angles)
    return torch.tensor(anchors, dtype=torch.float32)


def rotated_grid_anchors(grid_sizes: Sequence[Tuple[int, int]], strides: Sequence[int],
                         sizes: Sequence[Sequence[float]], aspect_ratios: Sequence[Sequence[float]],
                         angles: Sequence[Sequence[float]], offset: float = 0.0) -> List[torch.Tensor]:
    n = len(grid_sizes)
    if len(sizes) == 1:
        sizes = list(sizes) * n
    if len(aspect_ratios) == 1:
        aspect_ratios = list(aspect_ratios) * n
    if len(angles) == 1:
        angles = list(angles) * n
    out = []
    for (gh, gw), stride, s, ar, an in zip(grid_sizes, strides, sizes, aspect_ratios, angles):
        base = rotated_cell_anchors(s, ar, an)
        shifts_x = torch.arange(offset * stride, gw * stride, step=stride, dtype=torch.float32)
        shifts_y = torch.arange(offset * stride, gh * stride, step=stride, dtype=torch.float32)
        shift_y, shift_x = torch.meshgrid(shifts_y, shifts_x, indexing="ij")
        shift_x = shift_x.reshape(-1)
        shift_y = shift_y.reshape(-1)
        zeros = torch.zeros_like(shift_x)
        shifts = torch.stack((shift_x, shift_y, zeros, zeros, zeros), dim=1)
        out.append((shifts.view(-1, 1, 5) + base.view(1, -1, 5)).reshape(-1, 5))
    return out


# --------------------------------------------------------------------------------------
# find_top_rrpn_proposals (d2 modeling/proposal_generator/rrpn.py) -- A.4
# --------------------------------------------------------------------------------------
def find_top_rrpn_proposals(proposals: List[torch.Tensor], logits: List[torch.Tensor],
                            image_sizes: Sequence[Tuple[int, int]], nms_thresh: float,
                            pre_nms_topk: int, post_nms_topk: int, min_box_size: float = 0.0,
                            taps: dict = None):
    """proposals[l]: [N, HWA, 5]; logits[l]: [N, HWA].  Returns per image (boxes, logits)."""
    num_images = len(image_sizes)
    topk_scores, topk_proposals, level_ids = [], [], []
    batch_idx = torch.arange(num_images)
    for level_id, (proposals_i, logits_i) in enumerate(zip(proposals, logits)):
        hwa = logits_i.shape[1]
        k = min(pre_nms_topk, hwa)
        logits_sorted, idx = logits_i.sort(descending=True, dim=1, stable=True)
        topk_scores_i = logits_sorted[batch_idx, :k]
        topk_idx = idx[batch_idx, :k]
        topk_proposals_i = proposals_i[batch_idx[:, None], topk_idx]
        topk_proposals.append(topk_proposals_i)
        topk_scores.append(topk_scores_i)
        level_ids.append(torch.full((k,), level_id, dtype=torch.int64))
    topk_scores = torch.cat(topk_scores, dim=1)
    topk_proposals = torch.cat(topk_proposals, dim=1)
    level_ids = torch.cat(level_ids, dim=0)
    if taps is not None:
        taps["rpn_topk_boxes"] = topk_proposals.clone()
        taps["rpn_topk_scores"] = topk_scores.clone()
    results = []
    for n, image_size in enumerate(image_sizes):
        boxes = topk_proposals[n].clone()
        scores = topk_scores[n]
        lvl = level_ids
        valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores)
        if not valid.all():
            boxes, scores, lvl = boxes[valid], scores[valid], lvl[valid]
        clip_rotated_(boxes, image_size)
        keep = nonempty_rotated(boxes, min_box_size)
        if keep.sum().item() != len(boxes):
            boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
        keep = batched_nms_rotated(boxes, scores, lvl, nms_thresh)
        keep = keep[:post_nms_topk]
        results.append((boxes[keep], scores[keep]))
    return results


# --------------------------------------------------------------------------------------
# ROIPooler (d2 modeling/poolers.py) -- A.5
#   built at recognizers_hybrid_head.py:200-205 / :464-469 / :495-500
# --------------------------------------------------------------------------------------
def assign_boxes_to_levels(boxes: torch.Tensor, min_level: int, max_level: int,
                           canonical_box_size: int = 224, canonical_level: int = 4) -> torch.Tensor:
    box_sizes = torch.sqrt(boxes[:, 2] * boxes[:, 3])
    lvl = torch.floor(canonical_level + torch.log2(box_sizes / canonical_box_size + 1e-8))
    lvl = torch.clamp(lvl, min=min_level, max=max_level)
    return lvl.to(torch.int64) - min_level


def roi_pooler(features: List[torch.Tensor], box_lists: List[torch.Tensor], output_size,
               scales: Sequence[float], sampling_ratio: int) -> torch.Tensor:
    """features[l]: [N,C,H_l,W_l]; box_lists[i]: [M_i,5] rotated boxes of image i."""
    if isinstance(output_size, int):
        output_size = (output_size, output_size)
    min_level = -math.log2(scales[0])
    max_level = -math.log2(scales[-1])
    assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level))
    min_level, max_level = int(min_level), int(max_level)
    rois = torch.cat([torch.cat((torch.full((b.shape[0], 1), float(i)), b.float()), dim=1)
                      for i, b in enumerate(box_lists)], dim=0)
    num_levels = len(scales)
    if num_levels == 1:
        return roi_align_rotated(features[0], rois, output_size, scales[0], sampling_ratio)
    lvl = assign_boxes_to_levels(rois[:, 1:], min_level, max_level)
    m = rois.shape[0]
    out = torch.zeros((m, features[0].shape[1], output_size[0], output_size[1]), dtype=torch.float32)
    for level in range(num_levels):
        inds = torch.nonzero(lvl == level).squeeze(1)
        if inds.numel() == 0:
            continue
        out[inds] = roi_align_rotated(features[level], rois[inds], output_size, scales[level], sampling_ratio)
    return out
