"""Generate tests/golden/meta_postprocess.pt with the REFERENCE's own ``GlassRCNN._postprocess``
(glass/modeling/meta_arch/glass_rcnn.py:103-128), its ``PostProcessorRotatedBoxes.filter_small_boxes``
(post_processor_rotated_boxes.py:89-94), ``PostProcessorAcademic.resize_boxes`` (post_processor_academic.py:36-63) and
``detector_postprocess`` (:118-178).

Authoring container only.  detectron2 is not installable offline: ``GeneralizedRCNN`` is an empty base class (only the
unbound ``_postprocess`` is called, with a namespace as ``self``), ``Instances`` is this repo's field bag and
``RotatedBoxes.scale/clip/nonempty`` are the oracle's restatements of detectron2's (oracle/d2_ops.py).  The sequence,
the options, the in-place aliasing of ``pred_rboxes`` and the final indexing are the reference's code, unmodified.

    python tools/make_golden_meta_postprocess.py
"""
import contextlib
import importlib.util
import io
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
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
        def __getitem__(self, item):
            return RotatedBoxes(self.tensor[item].reshape(-1, 5))

        def clip(self, box_size, clip_angle_threshold=1.0):
            d2_ops.clip_rotated_(self.tensor, box_size, clip_angle_threshold)

        def scale(self, sx, sy):
            d2_ops.scale_rotated_(self.tensor, sx, sy)

        def nonempty(self, threshold=0.0):
            return d2_ops.nonempty_rotated(self.tensor, threshold)

    class Registry(dict):
        def __init__(self, name):
            super().__init__()

        def register(self, obj=None):
            def deco(o):
                self[o.__name__] = o
                return o
            return deco(obj) if obj is not None else deco

    class GeneralizedRCNN:
        pass

    stub("detectron2")
    stub("detectron2.config", configurable=lambda f=None, **kw: f)
    stub("detectron2.layers")
    stub("detectron2.layers.nms", nms_rotated=lambda b, s, iou_threshold: d2_ops.nms_rotated(b, s, iou_threshold))
    stub("detectron2.modeling")
    stub("detectron2.modeling.meta_arch")
    stub("detectron2.modeling.meta_arch.build", META_ARCH_REGISTRY=Registry("META_ARCH"))
    stub("detectron2.modeling.meta_arch.rcnn", GeneralizedRCNN=GeneralizedRCNN)
    stub("detectron2.structures", Instances=structures.Instances)
    stub("detectron2.structures.instances", Instances=structures.Instances)
    stub("detectron2.structures.boxes", BoxMode=object, Boxes=object, pairwise_ioa=None, pairwise_intersection=None)
    stub("detectron2.structures.rotated_boxes", pairwise_iou_rotated=lambda a, b: d2_ops.box_iou_rotated(a, b))
    stub("detectron2.utils")
    stub("detectron2.utils.memory", retry_if_cuda_oom=lambda f: f)
    stub("detectron2.utils.registry", Registry=Registry)
    for name in ("glass", "glass.postprocess", "glass.structures", "glass.modeling", "glass.modeling.recognition",
                 "glass.modeling.meta_arch", "glass.evaluation"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    stub("glass.modeling.recognition.text_encoder", TextEncoder=object)
    stub("glass.evaluation.text_evaluator", get_instances_text=None)

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m
    load("glass.structures.boxes", "glass/structures/boxes.py")
    rb = load("glass.postprocess.post_processor_rotated_boxes", "glass/postprocess/post_processor_rotated_boxes.py")
    sys.modules["glass.postprocess"].POST_PROCESSOR_REGISTRY = rb.POST_PROCESSOR_REGISTRY
    sys.modules["glass.postprocess"].build_post_processor = rb.build_post_processor
    ac = load("glass.postprocess.post_processor_academic", "glass/postprocess/post_processor_academic.py")
    meta = load("glass.modeling.meta_arch.glass_rcnn", "glass/modeling/meta_arch/glass_rcnn.py")
    return meta, ac, rb, RotatedBoxes, structures.Instances


def main():
    from golden_common import make_meta_postprocess_inputs
    meta, ac, rb, RotatedBoxes, Instances = load_reference()

    class PostProcessor:   # the two methods _postprocess calls on self.post_processor, bound to the reference's code
        min_box_dim = 2
        filter_small_boxes = rb.PostProcessorRotatedBoxes.filter_small_boxes
        resize_boxes = staticmethod(ac.PostProcessorAcademic.resize_boxes)

    cases = []
    specs = [  # seed, n, image (h, w), requested output (h, w) or None, MIN_BOX_DIMENSION, INFLATE_RATIO
        (0, 40, (800, 1000), None, None, None),          # GeneralizedRCNN behaviour (pretrain config)
        (1, 40, (800, 1000), None, 2, None),             # the three fine-tune configs
        (2, 64, (1216, 1216), (1024, 1024), 2, None),    # GlassRunner geometry: 1200/1216 canvas back to 1024
        (3, 48, (640, 960), (480, 1280), 2, 0.05),       # anisotropic output + inflation
        (4, 32, (512, 512), (768, 768), None, 0.1),
        (5, 0, (256, 256), None, 2, 0.05),               # empty
        (6, 8, (300, 300), (300, 300), 5, None),         # a larger limit
    ]
    for seed, n, hw, out_hw, min_dim, inflate in specs:
        boxes, scores = make_meta_postprocess_inputs(seed, n, hw)
        inst = Instances(hw, pred_boxes=RotatedBoxes(boxes.clone()), scores=scores.clone(), orig_idx=torch.arange(n))
        PostProcessor.min_box_dim = min_dim
        self_ns = types.SimpleNamespace(post_processor=PostProcessor(), filter_small_boxes=min_dim, inflate_ratio=inflate,
                                        drop_overlapping_boxes=None, ioa_threshold=None, valid_score=0)
        inp = {} if out_hw is None else {"height": out_hw[0], "width": out_hw[1]}
        with contextlib.redirect_stdout(io.StringIO()):   # resize_boxes prints a debug word (:38)
            out = meta.GlassRCNN._postprocess(self_ns, [inst], [inp], [hw])[0]["instances"]
        cases.append({"seed": seed, "n": n, "hw": hw, "out_hw": out_hw, "min_box_dim": min_dim, "inflate_ratio": inflate,
                      "idx": out.orig_idx.clone(), "boxes": out.pred_boxes.tensor.clone(), "scores": out.scores.clone(),
                      "image_size": tuple(out.image_size)})
        print(f"case {seed}: {n} -> {len(out)} boxes, output size {tuple(out.image_size)}")

    # detector_postprocess with pred_rboxes aliasing pred_boxes (forward_with_given_boxes under MASK_INFERENCE,
    # recognizers_hybrid_head.py:596-597) and with proposal_boxes only
    boxes, scores = make_meta_postprocess_inputs(20, 24, (400, 600))
    pb = RotatedBoxes(boxes.clone())
    inst = Instances((400, 600), pred_boxes=pb, scores=scores.clone(), orig_idx=torch.arange(24))
    inst.pred_rboxes = inst.pred_boxes
    out = ac.detector_postprocess(inst, 600, 800)
    alias = {"seed": 20, "n": 24, "hw": (400, 600), "out_hw": (600, 800), "idx": out.orig_idx.clone(),
             "boxes": out.pred_boxes.tensor.clone(), "rboxes": out.pred_rboxes.tensor.clone()}
    inst = Instances((400, 600), proposal_boxes=RotatedBoxes(boxes.clone()), objectness_logits=scores.clone(),
                     orig_idx=torch.arange(24))
    out = ac.detector_postprocess(inst, 200, 300)
    prop = {"seed": 20, "n": 24, "hw": (400, 600), "out_hw": (200, 300), "idx": out.orig_idx.clone(),
            "boxes": out.proposal_boxes.tensor.clone()}
    torch.save({"cases": cases, "alias": alias, "proposals": prop},
               os.path.join(ROOT, "tests", "golden", "meta_postprocess.pt"))
    print("saved", len(cases), "cases + alias + proposals")


if __name__ == "__main__":
    main()
