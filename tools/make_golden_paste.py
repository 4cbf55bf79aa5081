"""Generate tests/golden/paste_masks.pt with the REFERENCE's own rotated mask paste
(glass/postprocess/post_processor_academic.py:187-335: paste_masks_in_image / _do_paste_mask, the 5-column branch).
Authoring container only; the module's detectron2 / evaluator imports are stubbed (none is reached by the paste).

    python tools/make_golden_paste.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    class Registry(dict):
        def __init__(self, name):
            super().__init__()

        def register(self, obj=None):
            def deco(o):
                self[o.__name__] = o
                return o
            return deco(obj) if obj is not None else deco

    stub("detectron2")
    stub("detectron2.layers")
    stub("detectron2.layers.nms", nms_rotated=None)
    stub("detectron2.structures")
    stub("detectron2.structures.instances", Instances=object)
    stub("detectron2.structures.boxes", BoxMode=object, Boxes=object, pairwise_ioa=None, pairwise_intersection=None)
    stub("detectron2.structures.rotated_boxes", pairwise_iou_rotated=None)
    stub("detectron2.utils")
    stub("detectron2.utils.memory", retry_if_cuda_oom=lambda f: f)
    stub("detectron2.utils.registry", Registry=Registry)
    for name in ("glass", "glass.postprocess", "glass.structures", "glass.modeling", "glass.modeling.recognition",
                 "glass.evaluation"):
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
    return load("glass.postprocess.post_processor_academic", "glass/postprocess/post_processor_academic.py")


def main():
    from tests.golden_common import make_paste_inputs
    ac = load_reference()
    cases = []
    for seed, n, (h, w) in [(0, 6, (96, 128)), (1, 9, (130, 97)), (2, 1, (64, 64)), (3, 0, (48, 48))]:
        masks, boxes = make_paste_inputs(seed, n, h, w)
        out = ac.paste_masks_in_image(masks, boxes.clone(), (h, w), threshold=0.5)   # [n, h, w] bool
        soft, _ = ac._do_paste_mask(masks[:, None], boxes.clone(), h, w, skip_empty=False) if n else (torch.zeros(0, h, w), ())
        cases.append({"seed": seed, "n": n, "hw": (h, w),
                      "packed": torch.from_numpy(np.packbits(out.numpy().astype(np.uint8).reshape(-1))),
                      "count": int(out.sum()), "soft_sum": float(soft.double().sum()),
                      "soft_sample": soft[:, ::7, ::5].contiguous().clone()})
        print(f"case {seed}: {n} masks on {h}x{w}: {int(out.sum())} pixels set")
    torch.save({"cases": cases}, os.path.join(ROOT, "tests", "golden", "paste_masks.pt"))


if __name__ == "__main__":
    main()
