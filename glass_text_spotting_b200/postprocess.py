"""B200PostProcessor: the reference's ``PostProcessorAcademic`` (registered in POST_PROCESSOR_REGISTRY,
glass/postprocess/post_processor_academic.py:18-34, on top of PostProcessorRotatedBoxes,
post_processor_rotated_boxes.py:32-184) with the merge loop on the device (``glass_postprocess_merge``).

Same contract: ``post(preds: Instances) -> Instances`` -- the survivors, with merged ``pred_boxes``, a new
``pred_polygons`` field [k,4,2] and every other field carried along.  ``batch(...)`` is the sync-free form used
right after the hot path on padded per-image device tensors."""
from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from .structures import Instances, RotatedBoxes


@dataclass
class PostProcessingConfig:
    """cfg.POST_PROCESSING (glass/config.py:176-214); the academic fine-tune configs keep these defaults."""
    SKIP_ALL: bool = False
    MIN_BOX_DIMENSION: float = 2
    MERGE_IOA_THRESH: float = 0.3
    PAIRS_HEIGHT_RATIO_THRESH: float = 0.35
    VALID_CONFIDENCE: float = 0.15
    DETECT_THRESHOLD: float = 0.25
    TEXT_THRESHOLD: float = 0.25
    MAX_ANGLE_DIFF: float = 15


class B200PostProcessor:
    minimal_ioa_thresh = 0.01  # post_processor_rotated_boxes.py:40

    def __init__(self, cfg: Optional[PostProcessingConfig] = None, text_filter: bool = True, stop_index: int = 1):
        self.cfg = cfg or PostProcessingConfig()
        # post_processor_rotated_boxes.py:60-61
        assert self.cfg.VALID_CONFIDENCE <= self.cfg.DETECT_THRESHOLD, \
            "Valid score threshold must be smaller than the other class thresholds, to prevent word-in-word  cases"
        self.text_filter, self.stop_index = text_filter, stop_index

    def batch(self, boxes: torch.Tensor, scores: torch.Tensor, counts: Optional[torch.Tensor] = None,
              text_scores: Optional[torch.Tensor] = None):
        """boxes [n_img, m, 5], scores [n_img, m], counts int32 [n_img], text_scores [n_img, m] (all on the device)
        -> padded {boxes, scores, polygons, index, count, iters}; nothing is copied to the host."""
        c = self.cfg
        return ops.postprocess_merge(boxes, scores, counts, text_scores if self.text_filter else None,
                                     min_box_dim=c.MIN_BOX_DIMENSION, valid_score=c.VALID_CONFIDENCE,
                                     detect_threshold=c.DETECT_THRESHOLD, text_threshold=c.TEXT_THRESHOLD,
                                     merge_ioa_thresh=c.MERGE_IOA_THRESH,
                                     pairs_height_ratio_thresh=c.PAIRS_HEIGHT_RATIO_THRESH,
                                     max_angle_diff=c.MAX_ANGLE_DIFF, minimal_ioa_thresh=self.minimal_ioa_thresh)

    def __call__(self, preds: Instances, scale_ratio=1, **kwargs) -> Instances:
        if self.cfg.SKIP_ALL:
            return preds
        n = len(preds)
        dev = preds.scores.device if n or preds.has("scores") else "cuda"
        if n == 0:
            out = preds[torch.zeros(0, dtype=torch.int64, device=dev)] if preds.has("scores") else preds
            out._fields["pred_polygons"] = torch.zeros((0, 4, 2), dtype=torch.float32, device=dev)
            return out
        assert n <= 128, "the device post-processor holds at most 128 detections per image"
        boxes = preds.pred_boxes.tensor.reshape(1, n, 5).float().contiguous()
        scores = preds.scores.reshape(1, n).float().contiguous()
        ts = None
        if self.text_filter and preds.has("pred_text_prob"):
            ts = ops.text_scores(preds.pred_text_prob.float().contiguous(), self.stop_index).reshape(1, n)
        r = self.batch(boxes, scores, None, ts)
        k = int(r["count"][0].item())
        idx = r["index"][0, :k].long()
        out = preds[idx]
        out._fields["pred_boxes"] = RotatedBoxes(r["boxes"][0, :k].clone())
        out._fields["pred_polygons"] = r["polygons"][0, :k].clone()
        return out
