"""Meta-architecture: the GeneralizedRCNN / GlassRCNN inference surface on B200 kernels.

Mirrors glass/modeling/meta_arch/glass_rcnn.py:57-101 (``GlassRCNN.inference``; the pretrain config uses
d2's GeneralizedRCNN whose inference is the same sequence): preprocess -> backbone -> proposal generator ->
roi_heads (box branch, then recognizer on the detected boxes) -> detector_postprocess.
``model(batched_inputs: list[dict]) -> list[{"instances": Instances}]`` with the reference's field names
(pred_boxes, scores, pred_classes, orientations, pred_text_prob).

Unlike the reference (one image per forward) a batch of images is run
together through the dense stages; NMS, the decoder's early break and level assignment stay per image.
"""
from typing import Dict, List, Optional

import torch

from .. import ops
from ..structures import ImageList, Instances, RotatedBoxes
from .backbone import B200ResNetFPN, PIXEL_MEAN, PIXEL_STD
from .mask_head import B200MaskHead
from .roi_heads import B200GlassROIHeads
from .rpn import B200RotatedRPN


class B200GlassRCNN:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", mode: int = ops.MODE_SPLIT,
                 pixel_mean=PIXEL_MEAN, pixel_std=PIXEL_STD, mask_inference: bool = False, **head_kwargs):
        self.device = device
        self.pixel_mean, self.pixel_std = tuple(pixel_mean), tuple(pixel_std)
        self.backbone = B200ResNetFPN(state_dict, device=device, mode=mode, pixel_mean=pixel_mean, pixel_std=pixel_std)
        self.proposal_generator = B200RotatedRPN(state_dict, device=device, mode=mode)
        self.roi_heads = B200GlassROIHeads(state_dict, device=device, mode=mode, pixel_mean=pixel_mean,
                                           pixel_std=pixel_std, **head_kwargs)
        # MODEL.ROI_MASK_HEAD.MASK_INFERENCE (recognizers_hybrid_head.py:595-601): off in every shipped config
        self.mask_head = B200MaskHead(state_dict, device=device, mode=mode) if mask_inference else None

    # ------------------------------------------------------------------ a1
    def preprocess_image(self, batched_inputs: List[dict]) -> ImageList:
        """RAW pixels, padded with the pixel mean: (x - mean)/std is fused into the consumers (stem im2col and
        the image pooler), and mean-valued padding normalises to exactly 0 like ImageList.from_tensors."""
        imgs = [x["image"].to(self.device, non_blocking=True).float() for x in batched_inputs]
        return ImageList.from_tensors(imgs, self.backbone.size_divisibility, pad_value=self.pixel_mean)

    # ------------------------------------------------------------------ dense + decision stages on device
    def detect(self, images: torch.Tensor, img_hw: torch.Tensor, taps: Optional[dict] = None):
        feats = self.backbone(images)
        pb, ps, pi, pc = self.proposal_generator(feats, img_hw)
        det = self.roi_heads.forward_box(feats, pb, pc, img_hw, taps)
        if taps is not None:
            taps.update(features=feats, proposal_boxes=pb, objectness_logits=ps, proposal_count=pc)
        return feats, det

    def recognize(self, images: torch.Tensor, feats, det, counts_host: List[int], taps: Optional[dict] = None):
        n = images.shape[0]
        rois = []
        for i, c in enumerate(counts_host):
            if c > 0:
                b = det["pred_boxes"][i, :c]
                rois.append(torch.cat((torch.full((c, 1), float(i), device=b.device), b), 1))
        starts = [0]
        for c in counts_host:
            starts.append(starts[-1] + c)
        word_start = torch.tensor(starts, dtype=torch.int32, device=images.device)
        rois_t = torch.cat(rois).contiguous() if rois else torch.zeros((0, 6), device=images.device)
        return self.roi_heads.forward_recognizer(images, tuple(images.shape[-2:]), feats, rois_t, word_start, n, taps), starts

    @torch.no_grad()
    def forward_device(self, images: torch.Tensor, img_hw: torch.Tensor, taps: Optional[dict] = None):
        """Whole hot path on device tensors (images: RAW fp32 [n,3,H,W] padded to /32 with the pixel mean).
        One host sync (the detection counts size the recognizer's batch)."""
        feats, det = self.detect(images, img_hw, taps)
        counts_host = det["count"].cpu().tolist()
        probs, starts = self.recognize(images, feats, det, counts_host, taps)
        return det, probs, counts_host, starts

    def pack_detections(self, det, probs: torch.Tensor, counts_host: List[int], starts: List[int]) -> torch.Tensor:
        """Fixed-size record per image for the end-of-loop all-gather:
        [n, max_det, 1 + 5 + 1 + 1 + 2 + steps*classes] = (valid, box, score, class, orientation, text probs)."""
        n, m = det["pred_boxes"].shape[0], det["pred_boxes"].shape[1]
        tp = self.roi_heads.steps * self.roi_heads.num_classes
        rec = torch.zeros((n, m, 10 + tp), dtype=torch.float32, device=probs.device)
        for i, c in enumerate(counts_host):
            if c == 0:
                continue
            rec[i, :c, 0] = 1.0
            rec[i, :c, 1:6] = det["pred_boxes"][i, :c]
            rec[i, :c, 6] = det["scores"][i, :c]
            rec[i, :c, 8:10] = det["orientations"][i, :c]
            rec[i, :c, 10:] = probs[starts[i]: starts[i + 1]].reshape(c, tp)
        return rec

    @torch.no_grad()
    def inference(self, batched_inputs: List[dict], detected_instances=None, do_postprocess: bool = True,
                  taps: Optional[dict] = None):
        assert detected_instances is None, "given-box inference is not on the benchmarked path"
        il = self.preprocess_image(batched_inputs)
        img_hw = torch.tensor(il.image_sizes, dtype=torch.float32, device=self.device)
        keep = {} if self.mask_head is not None and taps is None else taps
        det, probs, counts_host, starts = self.forward_device(il.tensor, img_hw, keep)
        masks = None
        if self.mask_head is not None:   # _forward_mask on the detected boxes (recognizers_hybrid_head.py:598-601)
            rois = [torch.cat((torch.full((c, 1), float(i), device=self.device), det["pred_boxes"][i, :c]), 1)
                    for i, c in enumerate(counts_host) if c > 0]
            rois_t = torch.cat(rois).contiguous() if rois else torch.zeros((0, 6), device=self.device)
            masks = self.mask_head(keep["features"], rois_t, cap=len(counts_host) * self.roi_heads.max_det)
        results = []
        for i, c in enumerate(counts_host):
            inst = Instances(il.image_sizes[i],
                             pred_boxes=RotatedBoxes(det["pred_boxes"][i, :c].clone()),
                             scores=det["scores"][i, :c], pred_classes=torch.zeros(c, dtype=torch.int64, device=self.device),
                             orientations=det["orientations"][i, :c], pred_text_prob=probs[starts[i]: starts[i + 1]])
            if masks is not None:
                inst.pred_masks = masks[starts[i]: starts[i + 1]]
            if do_postprocess:
                inp = batched_inputs[i]
                inst = detector_postprocess(inst, inp.get("height", il.image_sizes[i][0]), inp.get("width", il.image_sizes[i][1]))
            results.append({"instances": inst})
        return results

    def forward(self, batched_inputs: List[dict]):
        return self.inference(batched_inputs)

    __call__ = forward


def detector_postprocess(results: Instances, output_height: int, output_width: int) -> Instances:
    """d2 detector_postprocess (boxes only; glass/postprocess/post_processor_academic.py:118-178 without masks):
    rescale to the requested output size, clip, drop empty boxes."""
    sx = output_width / results.image_size[1]
    sy = output_height / results.image_size[0]
    out = Instances((output_height, output_width), **results.get_fields())
    boxes = RotatedBoxes(out.pred_boxes.tensor.clone())
    boxes.scale(sx, sy)
    boxes.clip((output_height, output_width))
    out._fields["pred_boxes"] = boxes
    out = out[boxes.nonempty()]
    if out.has("pred_masks"):   # paste_masks_in_image on the rescaled boxes (post_processor_academic.py:163-169)
        out._fields["pred_masks"] = B200MaskHead.paste(out.pred_masks, out.pred_boxes.tensor, (output_height, output_width), 0.5)
    return out
