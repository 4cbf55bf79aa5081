"""Generate tests/golden/box_inference.pt with the REFERENCE's own box-branch inference
(glass/modeling/roi_heads/rotated_fast_rcnn.py: RotatedFastRCNNOutputs.inference :344-373 -> fast_rcnn_inference ->
fast_rcnn_inference_single_image_rotated :88-148, predict_probs / predict_orientations :480-491).

Authoring container only.  detectron2 is not installable offline; its pieces that this code touches are stubbed with
the oracle's restatements (RotatedBoxes.clip, Box2BoxTransformRotated.apply_deltas, batched_nms_rotated -- pinned by
detectron2's upstream KATs, tests/test_oracle_d2_ops.py).  The filtering, ordering, orientation handling and top-k are
the reference's code, unmodified.

    python tools/make_golden_box_inference.py
"""
import importlib.util
import os
import sys
import types

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    from glass_text_spotting_b200 import structures
    from oracle import d2_ops

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

    def configurable(fn=None, **kw):
        return fn

    stub("fvcore")
    stub("fvcore.nn", smooth_l1_loss=None)
    stub("detectron2")
    stub("detectron2.utils")
    stub("detectron2.utils.events", get_event_storage=None)
    stub("detectron2.config", configurable=configurable)
    stub("detectron2.layers", Linear=nn.Linear, ShapeSpec=object, cat=torch.cat,
         nonzero_tuple=lambda x: x.nonzero().unbind(1),
         batched_nms_rotated=lambda b, s, i, t: d2_ops.batched_nms_rotated(b, s, i, t))
    stub("detectron2.modeling")
    stub("detectron2.modeling.box_regression", Box2BoxTransformRotated=Box2BoxTransformRotated)
    stub("detectron2.structures", Boxes=object, Instances=structures.Instances, RotatedBoxes=RotatedBoxes)
    for name in ("glass", "glass.modeling", "glass.modeling.roi_heads", "glass.modeling.losses", "glass.utils"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m
    load("glass.utils.rotated_box_utils", "glass/utils/rotated_box_utils.py")
    load("glass.modeling.losses.rotated_box_losses", "glass/modeling/losses/rotated_box_losses.py")
    mod = load("glass.modeling.roi_heads.rotated_fast_rcnn", "glass/modeling/roi_heads/rotated_fast_rcnn.py")
    return mod, RotatedBoxes, Box2BoxTransformRotated, structures.Instances


def main():
    from tests.golden_common import make_box_inference_inputs
    mod, RotatedBoxes, B2B, Instances = load_reference()
    cases = []
    for seed, r, hw in [(0, 100, (512, 640)), (1, 37, (300, 300)), (2, 100, (1024, 1024)), (3, 5, (64, 96))]:
        logits, deltas, orient, proposals = make_box_inference_inputs(seed, r, hw)
        prop = Instances(hw, proposal_boxes=RotatedBoxes(proposals.clone()))
        out = mod.RotatedFastRCNNOutputs(B2B((10.0, 10.0, 5.0, 5.0, 10.0)), logits, deltas, orient, [prop])
        insts, kept = out.inference(score_thresh=0.05, nms_thresh=0.35, topk_per_image=100)
        i = insts[0]
        cases.append({"seed": seed, "r": r, "hw": hw, "pred_boxes": i.pred_boxes.tensor.clone(), "scores": i.scores.clone(),
                      "pred_classes": i.pred_classes.clone(), "orientations": i.orientations.clone(), "kept": kept[0].clone()})
        print(f"case {seed}: {r} proposals -> {len(i)} detections")
    torch.save({"cases": cases}, os.path.join(ROOT, "tests", "golden", "box_inference.pt"))


if __name__ == "__main__":
    main()
