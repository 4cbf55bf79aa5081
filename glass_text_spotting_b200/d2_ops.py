"""The detectron2 operator surface of SURVEY.md 8(b), under detectron2's own names and semantics, on the C-ABI kernels.

The reference reaches detectron2's native operators at these call sites:

* ``ROIPooler(...)(x: list[Tensor NCHW], box_lists: list[RotatedBoxes])`` -- built at
  glass/modeling/fusion/recognizers_hybrid_head.py:200-205 (box), :464-469 (recognizer), :495-500 (image crops),
  run at :320, :550, :556; inside it ``torch.ops.detectron2.roi_align_rotated_forward``.
* ``nms_rotated(boxes, scores, iou_threshold)`` -- glass/postprocess/post_processor_rotated_boxes.py:181;
  ``batched_nms_rotated`` -- glass/modeling/roi_heads/rotated_fast_rcnn.py:131.
* ``pairwise_iou_rotated`` -> ``torch.ops.detectron2.box_iou_rotated`` -- glass/structures/boxes.py:34, wrapped by the
  reference's own ``pairwise_ioa_rotated`` (:24-49).

The fused hot path (modeling/) does not go through this module -- it feeds the same kernels split-fp16 NHWC activations
and keeps every decision on the device.  These wrappers are the drop-in for code that calls the operators directly
(the reference's post-processor, evaluation and user scripts): NCHW fp32 in, NCHW fp32 out, device tensors only.
There is no CPU path: a CPU tensor raises.
"""
import math
from typing import List, Sequence, Union

import torch

from . import lib as _lib
from . import ops
from .structures import RotatedBoxes


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("glass_text_spotting_b200.d2_ops: CUDA tensors only (there is no CPU fallback)")


def _boxes5(b: Union[torch.Tensor, RotatedBoxes]) -> torch.Tensor:
    t = b.tensor if hasattr(b, "tensor") else b
    assert t.dim() == 2 and t.shape[1] == 5, "rotated boxes are [n, 5] = (cx, cy, w, h, angle_deg)"
    return t.float().contiguous()


# ------------------------------------------------------------------------------------------------- rotated IoU / IoA
def _pairwise(b1, b2, mode: int) -> torch.Tensor:
    b1, b2 = _boxes5(b1), _boxes5(b2)
    _need_cuda(b1, b2)
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    _lib.check(_lib.load().glass_box_iou_rotated(ops._ptr(b1), b1.shape[0], ops._ptr(b2), b2.shape[0], mode,
                                                 ops._ptr(out), ops._stream()))
    return out


def box_iou_rotated(boxes1: torch.Tensor, boxes2: torch.Tensor) -> torch.Tensor:
    """torch.ops.detectron2.box_iou_rotated: [n1,5] x [n2,5] -> IoU [n1,n2]."""
    return _pairwise(boxes1, boxes2, 0)


def pairwise_iou_rotated(boxes1: RotatedBoxes, boxes2: RotatedBoxes) -> torch.Tensor:
    """detectron2.structures.rotated_boxes.pairwise_iou_rotated (RotatedBoxes or tensors)."""
    return _pairwise(boxes1, boxes2, 0)


def pairwise_ioa_rotated(boxes1_tensor: torch.Tensor, boxes2_tensor: torch.Tensor) -> torch.Tensor:
    """glass.structures.boxes.pairwise_ioa_rotated (glass/structures/boxes.py:24-49): intersection over the smaller
    area.  Like the reference it takes TENSORS (it reads ``.shape[1]``)."""
    assert (boxes1_tensor.shape[1] == 5) and (boxes2_tensor.shape[1] == 5), "Input tensors don't describe rotated boxes"
    return _pairwise(boxes1_tensor, boxes2_tensor, 1)


# ------------------------------------------------------------------------------------------------- rotated NMS
def nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """torch.ops.detectron2.nms_rotated / detectron2.layers.nms_rotated: indices (int64) of the boxes kept by greedy
    rotated NMS, in descending-score order.  Up to 8192 boxes; the count is the op's only host read (it sizes the
    result, as it does in detectron2)."""
    boxes = _boxes5(boxes)
    _need_cuda(boxes, scores)
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    order = torch.sort(scores.float(), descending=True, stable=True).indices.contiguous()
    L = _lib.load()
    ws = torch.empty((L.glass_nms_rotated_all_workspace_bytes(n),), dtype=torch.uint8, device=boxes.device)
    keep = torch.empty((n,), dtype=torch.int64, device=boxes.device)
    count = torch.empty((1,), dtype=torch.int32, device=boxes.device)
    _lib.check(L.glass_nms_rotated_all(ops._ptr(boxes), ops._ptr(order), n, float(iou_threshold), ops._ptr(keep),
                                       ops._ptr(count), ops._ptr(ws), ws.numel(), ops._stream()))
    return keep[: int(count.item())]


def batched_nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """detectron2.layers.batched_nms_rotated: NMS within each category, by shifting every category's boxes into its own
    region of the plane (the offset is ``idx * (max_coordinate - min_coordinate + 1)``, as in detectron2 v0.6)."""
    assert boxes.shape[-1] == 5
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    boxes = boxes.float()
    max_coordinate = (torch.max(boxes[:, 0], boxes[:, 1]) + torch.max(boxes[:, 2], boxes[:, 3]) / 2).max()
    min_coordinate = (torch.min(boxes[:, 0], boxes[:, 1]) - torch.max(boxes[:, 2], boxes[:, 3]) / 2).min()
    offsets = idxs.to(boxes) * (max_coordinate - min_coordinate + 1)
    boxes_for_nms = boxes.clone()
    boxes_for_nms[:, :2] += offsets[:, None]
    return nms_rotated(boxes_for_nms, scores, iou_threshold)


# ------------------------------------------------------------------------------------------------- rotated RoIAlign
def _f32map(x: torch.Tensor) -> "ops.F32Map":
    """NCHW fp32 -> the kernel's padded fp32 NHWC map (border 1, channels padded to a multiple of 4)."""
    n, c, h, w = x.shape
    f = ops.F32Map(n, c, h, w, border=1, ld=ops.round_up(c, 4), device=x.device)
    f.buf[:, 1:-1, 1:-1, :c] = x.permute(0, 2, 3, 1)
    return f


def _roi_align_levels(xs: Sequence[torch.Tensor], rois: torch.Tensor, scales: Sequence[float], output_size, sampling_ratio,
                      min_level: int) -> torch.Tensor:
    """Channel blocks of <= 256 (the kernel's limit) through glass_roi_align_rotated; NHWC result -> NCHW view."""
    c = xs[0].shape[1]
    outs = []
    for c0 in range(0, c, 256):
        maps = [_f32map(x[:, c0: c0 + 256].float()) for x in xs]
        if maps[0].c % 4:   # pad channels are zero; the kernel wants a multiple of 4
            for m in maps:
                m.c = m.ld
        o = ops.roi_align_rotated(maps, rois, tuple(output_size), list(scales), sampling_ratio, min_level=min_level)
        outs.append(o[..., : min(256, c - c0)])
    out = outs[0] if len(outs) == 1 else torch.cat(outs, -1)
    return out.permute(0, 3, 1, 2)


def roi_align_rotated_forward(input: torch.Tensor, rois: torch.Tensor, spatial_scale: float, pooled_height: int,
                              pooled_width: int, sampling_ratio: int) -> torch.Tensor:
    """torch.ops.detectron2.roi_align_rotated_forward: input fp32 [N,C,H,W], rois [M,6] = (batch, cx, cy, w, h, deg)
    -> [M, C, pooled_height, pooled_width].  ``sampling_ratio`` 0 = adaptive grid ceil(roi / bins)."""
    _need_cuda(input, rois)
    assert input.dim() == 4 and rois.dim() == 2 and rois.shape[1] == 6
    if rois.shape[0] == 0:
        return torch.zeros((0, input.shape[1], pooled_height, pooled_width), dtype=torch.float32, device=input.device)
    return _roi_align_levels([input], rois.float().contiguous(), [spatial_scale], (pooled_height, pooled_width),
                             sampling_ratio, min_level=0)


class ROIAlignRotated:
    """detectron2.layers.ROIAlignRotated(output_size, spatial_scale, sampling_ratio)."""

    def __init__(self, output_size, spatial_scale: float, sampling_ratio: int):
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        self.spatial_scale, self.sampling_ratio = spatial_scale, sampling_ratio

    def forward(self, input: torch.Tensor, rois: torch.Tensor) -> torch.Tensor:
        return roi_align_rotated_forward(input, rois, self.spatial_scale, self.output_size[0], self.output_size[1],
                                         self.sampling_ratio)

    __call__ = forward


def convert_boxes_to_pooler_format(box_lists: List[RotatedBoxes]) -> torch.Tensor:
    """detectron2.modeling.poolers.convert_boxes_to_pooler_format: -> [M, 6] with the batch index in column 0."""
    ts = [_boxes5(b) for b in box_lists]
    idx = [torch.full((len(t), 1), float(i), dtype=torch.float32, device=t.device) for i, t in enumerate(ts)]
    return torch.cat([torch.cat((i, t), 1) for i, t in zip(idx, ts)], 0) if ts else torch.zeros((0, 6))


class ROIPooler:
    """detectron2.modeling.poolers.ROIPooler for ``pooler_type="ROIAlignRotated"`` (the only type the reference's
    rotated heads configure): ``scales`` are the strides' reciprocals of consecutive FPN levels; a box goes to level
    ``floor(canonical_level + log2(sqrt(w*h) / canonical_box_size + 1e-8))`` clamped to the available ones.  With
    several levels the assignment runs inside ONE kernel launch over all of them (no per-level index_put)."""

    def __init__(self, output_size, scales: Sequence[float], sampling_ratio: int, pooler_type: str = "ROIAlignRotated",
                 canonical_box_size: int = 224, canonical_level: int = 4):
        if pooler_type != "ROIAlignRotated":
            raise ValueError(f"Unknown pooler type: {pooler_type}")
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        assert len(self.output_size) == 2
        self.scales, self.sampling_ratio = [float(s) for s in scales], sampling_ratio
        min_level, max_level = -math.log2(self.scales[0]), -math.log2(self.scales[-1])
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level)), \
            "Featuremap stride is not power of 2!"
        self.min_level, self.max_level = int(min_level), int(max_level)
        assert len(self.scales) == self.max_level - self.min_level + 1, "[ROIPooler] Sizes of input featuremaps do not form a pyramid!"
        assert canonical_box_size > 0
        if (canonical_box_size, canonical_level) != (224, 4):
            raise NotImplementedError("the fused level assignment implements detectron2's defaults (224, 4), which are "
                                      "the only values the reference uses")

    def forward(self, x: List[torch.Tensor], box_lists: List[RotatedBoxes]) -> torch.Tensor:
        assert isinstance(x, list) and isinstance(box_lists, list), "Arguments to pooler must be lists"
        assert len(x) == len(self.scales), f"unequal value, num_level_assignments={len(self.scales)}, but x is list of {len(x)} Tensors"
        assert len(box_lists) == x[0].size(0), f"unequal value, x[0] batch dim 0 is {x[0].size(0)}, but box_list has length {len(box_lists)}"
        _need_cuda(*x)
        c = x[0].shape[1]
        if len(box_lists) == 0 or sum(len(b) for b in box_lists) == 0:
            return torch.zeros((0, c) + self.output_size, dtype=torch.float32, device=x[0].device)
        rois = convert_boxes_to_pooler_format(box_lists).to(x[0].device)
        return _roi_align_levels(x, rois, self.scales, self.output_size, self.sampling_ratio, min_level=self.min_level)

    __call__ = forward


__all__ = ["box_iou_rotated", "pairwise_iou_rotated", "pairwise_ioa_rotated", "nms_rotated", "batched_nms_rotated",
           "roi_align_rotated_forward", "ROIAlignRotated", "ROIPooler", "convert_boxes_to_pooler_format"]

