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


# where P2P3Fusion is forked onto a side stream: 0 = not at all (default: a same-box A/B on B200 showed no gain --
# 167.5 / 167.5 / 167.7 images/s -- the persistent GEMMs do not slip in beside the latency-bound detector kernels),
# 1 = after the backbone, 2 = after the RPN head's convs
_FORK = int(__import__("os").environ.get("GLASS_FORK", "0"))


class B200GlassRCNN:
    """``filter_small_boxes`` / ``inflate_ratio`` are GlassRCNN's ``_postprocess`` options
    (glass_rcnn.py:43-50: cfg.POST_PROCESSING.MIN_BOX_DIMENSION / INFLATE_RATIO; the three fine-tune configs set
    MIN_BOX_DIMENSION 2, none sets INFLATE_RATIO).  With both None the model is d2's plain GeneralizedRCNN, which is
    what configs/glass_pretrain.yaml:40 selects.  ``drop_overlapping_boxes`` (cfg.POST_PROCESSING.DROP_OVERLAPPING,
    in no shipped config) cannot run in the reference either: post_processor_academic.py:73 hands RotatedBoxes
    OBJECTS to pairwise_ioa_rotated, whose first statement reads ``.shape`` (glass/structures/boxes.py:31) and
    raises AttributeError -- requesting it here raises at construction instead of at the first image."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", mode: int = ops.MODE_SPLIT,
                 pixel_mean=PIXEL_MEAN, pixel_std=PIXEL_STD, mask_inference: bool = False,
                 filter_small_boxes: Optional[float] = None, inflate_ratio: Optional[float] = None,
                 drop_overlapping_boxes=None, rpn_kwargs: Optional[dict] = None, **head_kwargs):
        if drop_overlapping_boxes:
            raise NotImplementedError("POST_PROCESSING.DROP_OVERLAPPING raises AttributeError in the reference "
                                      "(post_processor_academic.py:73 -> glass/structures/boxes.py:31); not provided")
        self.device = device
        self.filter_small_boxes, self.inflate_ratio = filter_small_boxes, inflate_ratio
        self.pixel_mean, self.pixel_std = tuple(pixel_mean), tuple(pixel_std)
        self.backbone = B200ResNetFPN(state_dict, device=device, mode=mode, pixel_mean=pixel_mean, pixel_std=pixel_std)
        self.proposal_generator = B200RotatedRPN(state_dict, device=device, mode=mode, **(rpn_kwargs or {}))
        self.roi_heads = B200GlassROIHeads(state_dict, device=device, mode=mode, pixel_mean=pixel_mean,
                                           pixel_std=pixel_std, **head_kwargs)
        # MODEL.ROI_MASK_HEAD.MASK_INFERENCE (recognizers_hybrid_head.py:595-601): off in every shipped config
        self.mask_head = B200MaskHead(state_dict, device=device, mode=mode) if mask_inference else None
        self.roi_heads.mask_head = self.mask_head      # forward_with_given_boxes runs it (:595-601)

    # ------------------------------------------------------------------ a1
    def preprocess_image(self, batched_inputs: List[dict]) -> ImageList:
        """RAW pixels, padded with the pixel mean: (x - mean)/std is fused into the consumers (stem im2col and
        the image pooler), and mean-valued padding normalises to exactly 0 like ImageList.from_tensors."""
        imgs = [x["image"].to(self.device, non_blocking=True).float() for x in batched_inputs]
        return ImageList.from_tensors(imgs, self.backbone.size_divisibility, pad_value=self.pixel_mean)

    # ------------------------------------------------------------------ dense + decision stages on device
    _side = None

    def detect(self, images: torch.Tensor, img_hw: torch.Tensor, taps: Optional[dict] = None):
        feats = self.backbone(images)
        return feats, self.detect_from_features(feats, img_hw, taps)

    def detect_from_features(self, feats, img_hw: torch.Tensor, taps: Optional[dict] = None):
        pb, ps, pi, pc = self.proposal_generator.forward_device(feats, img_hw)
        det = self.roi_heads.forward_box(feats, pb, pc, img_hw, taps)
        if taps is not None:
            taps.update(features=feats, proposal_boxes=pb, objectness_logits=ps, proposal_count=pc)
        return det

    @torch.no_grad()
    def forward_packed(self, images: torch.Tensor, img_hw: torch.Tensor, taps: Optional[dict] = None):
        """The whole hot path with NO host synchronisation (images: RAW fp32 [n,3,H,W] padded to /32 with the pixel mean):
        backbone -> RPN -> box branch -> glass_pack_rois (detections become the recognizer's RoI list on the device) ->
        recognizer sized for the capacity n * max_det, every kernel reading the live word count from device memory ->
        glass_pack_detections.  Returns (rec [n, max_det, 10 + steps*classes] -- the fixed-size per-image record of the
        end-of-loop all-gather, SURVEY.md 8e --, det dict, probs [capacity, steps, classes], word_start int32 [n+1]);
        all device tensors in persistent workspaces, valid until the next call.  CUDA-graph capturable."""
        heads = self.roi_heads
        # P2P3Fusion (two full-GPU 1x1 GEMMs) depends only on the pyramid: forked onto a side stream, it fills the SMs
        # that the latency-bound middle of the detector (per-level top-k, the two rotated NMS passes, the 400-row FC head)
        # leaves idle; joined before the recognizer's pooler reads it.  Capturable: the fork / join become graph edges.
        feats = self.backbone(images)
        rpn = self.proposal_generator
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        fork = _FORK

        def side_p2p3():
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                return heads.p2p3(feats)
        gmap = side_p2p3() if fork == 1 else None
        preds = rpn.head(feats)
        if fork == 2:
            gmap = side_p2p3()
        boxes, scores = rpn.topk_decode(preds)
        pb, ps, pi, pc = rpn.select(boxes, scores, img_hw)
        det = heads.forward_box(feats, pb, pc, img_hw, taps)
        if taps is not None:
            taps.update(features=feats, proposal_boxes=pb, objectness_logits=ps, proposal_count=pc)
        if gmap is not None:
            cur.wait_stream(self._side)
        n, m = det["pred_boxes"].shape[0], det["pred_boxes"].shape[1]
        ws = heads.ws
        rois = ws.raw("step.rois", (n * m, 6), torch.float32)
        word_start = ws.raw("step.word_start", (n + 1,), torch.int32)
        total = ws.raw("step.total", (1,), torch.int32)
        ops.pack_rois(det["pred_boxes"], det["count"], rois, word_start, total)
        probs = heads.forward_recognizer(images, tuple(images.shape[-2:]), feats, rois, word_start, n, taps, n_dev=total,
                                         gmap=gmap)
        rec = ws.raw("step.rec", (n, m, 10 + heads.steps * heads.num_classes), torch.float32)
        ops.pack_detections(det["pred_boxes"], det["scores"], det["orientations"] if heads.orientation_on else None,
                            det["count"], probs, word_start, rec)
        return rec, det, probs, word_start

    @torch.no_grad()
    def forward_device(self, images: torch.Tensor, img_hw: torch.Tensor, taps: Optional[dict] = None):
        """``forward_packed`` + ONE read-back of the per-image detection counts AFTER everything is enqueued (the caller
        wants host-side Instances): returns (det, probs [K_total, steps, classes], counts_host, starts)."""
        _, det, probs, _ = self.forward_packed(images, img_hw, taps)
        counts_host = det["count"].cpu().tolist()
        starts = [0]
        for c in counts_host:
            starts.append(starts[-1] + c)
        return det, probs[: starts[-1]], counts_host, starts

    def pack_detections(self, det, probs: torch.Tensor, counts_host: List[int], starts: List[int]) -> torch.Tensor:
        """Fixed-size record per image for the end-of-loop all-gather:
        [n, max_det, 1 + 5 + 1 + 1 + 2 + steps*classes] = (valid, box, score, class, orientation, text probs);
        one kernel (glass_pack_detections).  ``forward_packed`` produces the same record without the host counts."""
        n, m = det["pred_boxes"].shape[0], det["pred_boxes"].shape[1]
        tp = self.roi_heads.steps * self.roi_heads.num_classes
        rec = torch.empty((n, m, 10 + tp), dtype=torch.float32, device=probs.device)
        word_start = torch.tensor(starts, dtype=torch.int32, device=probs.device)
        if probs.shape[0] == 0:
            probs = torch.zeros((1, self.roi_heads.steps, self.roi_heads.num_classes), device=rec.device)
        return ops.pack_detections(det["pred_boxes"], det["scores"],
                                   det["orientations"] if self.roi_heads.orientation_on else None, det["count"],
                                   probs.contiguous(), word_start, rec)

    # ------------------------------------------------------------------ the step as ONE CUDA graph
    @torch.no_grad()
    def graph_step(self, images: torch.Tensor, img_hw: torch.Tensor) -> torch.Tensor:
        """``forward_packed`` replayed from a CUDA graph (captured on first use per input shape, after two eager warm-up
        passes that allocate every workspace): one graph launch per step instead of ~190 kernel launches, no host work
        between them.  ``images`` (fp32, or uint8: converted by the copy) / ``img_hw`` are copied into the graph's static
        inputs; returns the static record buffer (valid until the next replay)."""
        key = (tuple(images.shape), images.device.index)
        g = self._graphs.get(key) if hasattr(self, "_graphs") else None
        if g is None:
            if not hasattr(self, "_graphs"):
                self._graphs = {}
            static_in = images.float().clone()
            static_hw = img_hw.clone()
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.forward_packed(static_in, static_hw)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            L = ops._lib.load()
            before = L.glass_launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                rec, det, probs, word_start = self.forward_packed(static_in, static_hw)
            g = self._graphs[key] = {"graph": graph, "in": static_in, "hw": static_hw, "rec": rec, "det": det,
                                     "probs": probs, "word_start": word_start,
                                     "launches": int(L.glass_launch_count() - before)}
        g["in"].copy_(images, non_blocking=True)
        g["hw"].copy_(img_hw, non_blocking=True)
        g["graph"].replay()
        self.last_graph = g
        return g["rec"]

    @torch.no_grad()
    def inference(self, batched_inputs: List[dict], detected_instances: Optional[List[Instances]] = None,
                  do_postprocess: bool = True, taps: Optional[dict] = None):
        """glass_rcnn.py:57-101.  ``detected_instances`` (one Instances per image with ``pred_boxes`` and
        ``pred_classes``) skips detection and only predicts the other per-RoI outputs (text, masks).
        Returns ``list[{"instances": Instances}]``, or the raw ``list[Instances]`` when ``do_postprocess`` is False."""
        il = self.preprocess_image(batched_inputs)
        if detected_instances is not None:
            feats = self.backbone(il.tensor)
            results = self.roi_heads.forward_with_given_boxes(il, feats, [x.to(self.device) for x in detected_instances])
            return self._postprocess(results, batched_inputs, il.image_sizes) if do_postprocess else results
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
                             scores=det["scores"][i, :c], pred_classes=torch.zeros(c, dtype=torch.int64, device=self.device))
            if self.roi_heads.orientation_on:    # MODEL.ORIENTATION_ON (rotated_fast_rcnn.py:141-142)
                inst.orientations = det["orientations"][i, :c]
            inst.pred_text_prob = probs[starts[i]: starts[i + 1]]
            if masks is not None:   # forward_with_given_boxes under MASK_INFERENCE (recognizers_hybrid_head.py:595-601)
                inst.pred_masks = masks[starts[i]: starts[i + 1]]
                inst.pred_rboxes = inst.pred_boxes      # :596-597 -- the SAME object (detector_postprocess relies on it)
            results.append(inst)
        return self._postprocess(results, batched_inputs, il.image_sizes) if do_postprocess else results

    def _postprocess(self, instances: List[Instances], batched_inputs: List[dict], image_sizes):
        """GlassRCNN._postprocess (glass_rcnn.py:103-128): optional small-box filter and inflation, then
        detector_postprocess to the requested output size.  Index glue on <= 100 boxes per image."""
        out = []
        for inst, inp, size in zip(instances, batched_inputs, image_sizes):
            height, width = inp.get("height", size[0]), inp.get("width", size[1])
            if self.filter_small_boxes:
                inst = filter_small_boxes(inst, self.filter_small_boxes)
            if self.inflate_ratio:
                inst = resize_boxes(inst, self.inflate_ratio)
            out.append({"instances": detector_postprocess(inst, height, width)})
        return out

    def forward(self, batched_inputs: List[dict]):
        return self.inference(batched_inputs)

    __call__ = forward


def filter_small_boxes(preds: Instances, min_box_dim: float) -> Instances:
    """PostProcessorRotatedBoxes.filter_small_boxes (post_processor_rotated_boxes.py:89-94): keep min(w, h) >= dim."""
    if len(preds) == 0:
        return preds
    boxes = preds.pred_boxes.tensor
    return preds[torch.min(boxes[:, 2], boxes[:, 3]) >= min_box_dim]


def resize_boxes(preds: Instances, ratio: float, axis: str = "both") -> Instances:
    """PostProcessorAcademic.resize_boxes (post_processor_academic.py:36-63): widen / heighten every box by
    ``ratio`` of its own width / height IN PLACE (like the reference), then clip to the image."""
    if len(preds) == 0:
        return preds
    if axis not in ("both", "vertical", "horizontal"):
        raise Exception('Please provide an axis value of either "both"/"horizontal"/"vertical')
    boxes = preds.pred_boxes.tensor
    delta_x = ratio * boxes[:, 2] if axis != "vertical" else 0
    delta_y = ratio * boxes[:, 3] if axis != "horizontal" else 0
    boxes[:, 2] += delta_x
    boxes[:, 3] += delta_y
    preds.pred_boxes.clip(preds.image_size)
    return preds


def detector_postprocess(results: Instances, output_height: int, output_width: int) -> Instances:
    """glass/postprocess/post_processor_academic.py:118-178: rescale ``pred_boxes`` (or ``proposal_boxes``) to the
    requested output size, clip, drop empty boxes, paste masks, rescale ``pred_rboxes``.  Unlike the reference the
    caller's box tensor is not modified in place (the result holds a scaled copy).  When ``pred_rboxes`` IS the
    ``pred_boxes`` object (forward_with_given_boxes under MASK_INFERENCE, recognizers_hybrid_head.py:596-597) the
    reference's in-place scale/clip of :158-159 has already moved it before :173-175 scale and clip it again; the
    result is reproduced (tests/golden/meta_postprocess.pt "alias")."""
    sx = output_width / results.image_size[1]
    sy = output_height / results.image_size[0]
    out = Instances((output_height, output_width), **results.get_fields())
    name = "pred_boxes" if out.has("pred_boxes") else "proposal_boxes"
    src = out.get(name)
    boxes = RotatedBoxes(src.tensor.clone())
    boxes.scale(sx, sy)
    boxes.clip((output_height, output_width))
    out._fields[name] = boxes
    if out.has("pred_rboxes"):
        rb = out.pred_rboxes
        out._fields["pred_rboxes"] = RotatedBoxes((boxes if rb is src else rb).tensor.clone())
    out = out[boxes.nonempty()]
    if out.has("pred_rboxes"):   # :173-175
        out.pred_rboxes.scale(sx, sy)
        out.pred_rboxes.clip((output_height, output_width))
    if out.has("pred_masks"):   # paste_masks_in_image on the rescaled boxes (post_processor_academic.py:163-169)
        out._fields["pred_masks"] = B200MaskHead.paste(out.pred_masks, out.pred_boxes.tensor, (output_height, output_width), 0.5)
    return out
