"""Generate tests/golden/box_branch.pt with the REFERENCE's own box branch, end to end:
``MaskRotatedRecognizerHybridHead._forward_box`` (glass/modeling/fusion/recognizers_hybrid_head.py:291-339) driving the
reference's own ``RotatedFastRCNNOutputLayers`` (glass/modeling/roi_heads/rotated_fast_rcnn.py:495-620), which the
reference's own ``from_config`` builds from the REFERENCE's configs/glass_pretrain.yaml, and its
``RotatedFastRCNNOutputs.inference`` (:344-373 -> :88-148).

This pins the WIRING of rows a5-a8: pooled features -> box head -> the three predictor Linears (names,
widths, flatten order) -> decode weights / score threshold / NMS threshold / top-k as the config supplies them -> the
result fields.  The inference arithmetic on its own is already pinned by tests/golden/box_inference.pt.

Authoring container only.  detectron2 is not installable offline; stubbed with published semantics or the oracle's
restatements (pinned by detectron2's upstream known-answer tests): ``configurable``, ``Linear`` (= nn.Linear), ``cat``,
``nonzero_tuple``, ``ROIPooler`` / ``batched_nms_rotated`` / ``Box2BoxTransformRotated.apply_deltas`` /
``RotatedBoxes.clip`` (oracle/d2_ops.py), ``FastRCNNConvFCHead`` (NUM_CONV 0, NUM_FC 2: flatten, fc1, ReLU, fc2, ReLU).
The orchestration and the predictor are the reference's code, unmodified.

    python tools/make_golden_box_branch.py
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class FastRCNNConvFCHead(nn.Module):
    """detectron2.modeling.roi_heads.box_head.FastRCNNConvFCHead for NUM_CONV 0 / NUM_FC 2."""

    def __init__(self, cin, dim):
        super().__init__()
        self.fc1 = nn.Linear(cin, dim)
        self.fc2 = nn.Linear(dim, dim)

    def forward(self, x):
        x = torch.flatten(x, start_dim=1)
        return F.relu(self.fc2(F.relu(self.fc1(x))))


def load_reference():
    from glass_text_spotting_b200 import structures
    from glass_text_spotting_b200.config import CfgNode
    from oracle import d2_ops

    class Registry(dict):
        def __init__(self, name):
            super().__init__()

        def register(self, obj=None):
            def deco(o):
                self[o.__name__] = o
                return o
            return deco(obj) if obj is not None else deco

        def get(self, name):
            return self[name]

    def configurable(init_func=None, *, from_config=None):
        def wrapped(self, *args, **kwargs):
            if args and isinstance(args[0], CfgNode):
                init_func(self, **type(self).from_config(*args, **kwargs))
            else:
                init_func(self, *args, **kwargs)
        return wrapped

    class RotatedBoxes(structures.RotatedBoxes):
        @classmethod
        def cat(cls, boxes_list):
            return cls(torch.cat([b.tensor for b in boxes_list], 0))

        def clip(self, box_size, clip_angle_threshold=1.0):
            d2_ops.clip_rotated_(self.tensor, box_size, clip_angle_threshold)

    class Box2BoxTransformRotated:
        def __init__(self, weights):
            self.weights = weights

        def apply_deltas(self, deltas, boxes):
            return d2_ops.apply_deltas_rotated(deltas, boxes, self.weights)

    class ROIPooler:
        def __init__(self, output_size, scales, sampling_ratio, pooler_type):
            assert pooler_type == "ROIAlignRotated"
            self.output_size, self.scales, self.sampling_ratio = output_size, list(scales), sampling_ratio

        def __call__(self, x, box_lists):
            return d2_ops.roi_pooler(x, [b.tensor for b in box_lists], self.output_size, self.scales, self.sampling_ratio)

    class ShapeSpec:
        def __init__(self, channels=None, height=None, width=None, stride=None):
            self.channels, self.height, self.width, self.stride = channels, height, width, stride

    class StandardROIHeads(nn.Module):
        pass

    wi = stub("fvcore.nn.weight_init", c2_msra_fill=lambda m: None, c2_xavier_fill=lambda m: None)
    stub("fvcore")
    stub("fvcore.nn", weight_init=wi, smooth_l1_loss=None)
    stub("detectron2")
    stub("detectron2.config", configurable=configurable)
    stub("detectron2.layers", Linear=nn.Linear, ShapeSpec=ShapeSpec, cat=torch.cat, Conv2d=None, get_norm=None,
         nonzero_tuple=lambda x: x.nonzero().unbind(1),
         batched_nms_rotated=lambda b, s, i, t: d2_ops.batched_nms_rotated(b, s, i, t))
    stub("detectron2.modeling")
    stub("detectron2.modeling.box_regression", Box2BoxTransformRotated=Box2BoxTransformRotated)
    stub("detectron2.modeling.poolers", ROIPooler=ROIPooler)
    stub("detectron2.modeling.roi_heads")
    stub("detectron2.modeling.roi_heads.box_head", build_box_head=None)
    stub("detectron2.modeling.roi_heads.mask_head", build_mask_head=None, ROI_MASK_HEAD_REGISTRY=Registry("ROI_MASK_HEAD"))
    stub("detectron2.modeling.roi_heads.roi_heads", ROI_HEADS_REGISTRY=Registry("ROI_HEADS"), StandardROIHeads=StandardROIHeads,
         select_foreground_proposals=None)
    stub("detectron2.modeling.roi_heads.rotated_fast_rcnn", RotatedFastRCNNOutputLayers=None)
    stub("detectron2.structures", ImageList=structures.ImageList, Instances=structures.Instances, RotatedBoxes=RotatedBoxes,
         pairwise_iou_rotated=None, Boxes=object)
    stub("detectron2.utils")
    stub("detectron2.utils.events", get_event_storage=None)
    stub("detectron2.utils.registry", Registry=Registry)
    for name in ("glass", "glass.modeling", "glass.modeling.fusion", "glass.modeling.recognition", "glass.modeling.roi_heads",
                 "glass.modeling.losses", "glass.structures", "glass.utils"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    # glass modules the hybrid head imports but the box branch never reaches
    stub("glass.modeling.fusion.fusion_modules", P2P3Fusion=None, build_hybrid_feature_fusion=None)
    stub("glass.modeling.fusion.local_feature_extraction", build_hybrid_feature_extractor=None)
    stub("glass.modeling.recognition.recognizer_head_v2", build_recognizer_head=None)
    stub("glass.modeling.recognition.recognizer_pooler_pad", build_recognizer_pooler_pad=None)
    stub("glass.modeling.roi_heads.rotated_head", add_ground_truth_to_proposals=None)
    stub("glass.structures.boxes", box_to_rbox=None, rbox_to_box=None)

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m
    load("glass.utils.rotated_box_utils", "glass/utils/rotated_box_utils.py")
    load("glass.modeling.losses.rotated_box_losses", "glass/modeling/losses/rotated_box_losses.py")
    fr = load("glass.modeling.roi_heads.rotated_fast_rcnn", "glass/modeling/roi_heads/rotated_fast_rcnn.py")
    hh = load("glass.modeling.fusion.recognizers_hybrid_head", "glass/modeling/fusion/recognizers_hybrid_head.py")
    return hh, fr, ShapeSpec, ROIPooler, RotatedBoxes, structures


def reference_config(detections: int):
    from glass_text_spotting_b200 import config
    cfg = config.load_config(os.path.join(REF, "configs/glass_pretrain.yaml"))
    cfg.merge({"TEST": {"DETECTIONS_PER_IMAGE": detections},
               "MODEL": {"ROI_BOX_HEAD": {"SMOOTH_L1_BETA": 1.0, "BBOX_REG_LOSS_TYPE": "sine_square_loss", "BBOX_REG_LOSS_WEIGHT": 1.0},
                         "ROI_ORIENTATION_HEAD": {"LOSS_WEIGHT": 0.3}}})
    return cfg


def main():
    from golden_common import make_box_branch_inputs, seeded_fill
    hh, fr, ShapeSpec, ROIPooler, RotatedBoxes, st = load_reference()
    cases = []
    with torch.no_grad():
        for seed, r, detections in [(0, 60, 100), (1, 100, 12), (2, 5, 100)]:
            cfg = reference_config(detections)
            bh = cfg.MODEL.ROI_BOX_HEAD
            ns = types.SimpleNamespace(
                training=False, box_in_features=list(cfg.MODEL.ROI_HEADS.IN_FEATURES),
                # recognizers_hybrid_head.py:184-217 (_init_box_head)
                box_pooler=ROIPooler(output_size=bh.POOLER_RESOLUTION, scales=[1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64],
                                     sampling_ratio=bh.POOLER_SAMPLING_RATIO, pooler_type=bh.POOLER_TYPE),
                box_head=FastRCNNConvFCHead(256 * bh.POOLER_RESOLUTION ** 2, bh.FC_DIM).eval(),
                box_predictor=fr.RotatedFastRCNNOutputLayers(cfg, ShapeSpec(channels=bh.FC_DIM)).eval())
            seeded_fill(ns.box_head, 700 + seed)
            seeded_fill(ns.box_predictor, 710 + seed)
            ns.box_predictor.cls_score.weight.mul_(4.0)      # spread the scores around the 0.05 threshold
            ns.box_predictor.bbox_pred.weight.mul_(0.3)      # keep the decoded boxes near their proposals (NMS bites)
            feats, proposals, hw = make_box_branch_inputs(seed, r)
            inst = st.Instances(hw, proposal_boxes=RotatedBoxes(proposals.clone()), objectness_logits=torch.zeros(r))
            out = hh.MaskRotatedRecognizerHybridHead._forward_box(ns, feats, [inst])[0]
            cases.append({"seed": seed, "r": r, "detections": detections, "hw": hw,
                          "head_keys": sorted(ns.box_head.state_dict()), "predictor_keys": sorted(ns.box_predictor.state_dict()),
                          "pred_boxes": out.pred_boxes.tensor.clone(), "scores": out.scores.clone(),
                          "pred_classes": out.pred_classes.clone(), "orientations": out.orientations.clone(),
                          "image_size": tuple(out.image_size),
                          "thresholds": (ns.box_predictor.test_score_thresh, ns.box_predictor.test_nms_thresh,
                                         ns.box_predictor.test_topk_per_image, tuple(ns.box_predictor.box2box_transform.weights))})
            print(f"case {seed}: {r} proposals -> {len(out)} detections (top-k {detections}), scores "
                  f"{float(out.scores.min()) if len(out) else 0:.3f}..{float(out.scores.max()) if len(out) else 0:.3f}")
    torch.save({"cases": cases}, os.path.join(ROOT, "tests", "golden", "box_branch.pt"))


if __name__ == "__main__":
    main()
