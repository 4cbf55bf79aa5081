"""Generate tests/golden/runner.pt with the REFERENCE's own ``GlassRunner`` methods (glass/inference/glass_runner.py):
``get_inference_scale_ratio`` (:111-121), ``_image_to_tensor`` (:123-148) and the ``__call__`` flow (:72-109) around a
recording stand-in for the model and the post-processor.  Authoring container only; the module's detectron2 / glass
imports are stubbed (none is reached by these methods), ``rgb2grey`` is the reference's (glass/utils/common_utils.py).

    python tools/make_golden_runner.py
"""
import importlib.util
import logging
import os
import sys
import types

import numpy as np
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
    stub("detectron2")
    stub("detectron2.checkpoint", DetectionCheckpointer=None)
    stub("detectron2.config", get_cfg=None)
    stub("detectron2.modeling", build_model=None)
    stub("detectron2.structures", Instances=structures.Instances)
    for name in ("glass", "glass.utils", "glass.modeling", "glass.modeling.recognition"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    stub("glass.config", add_e2e_config=None, add_glass_config=None, add_dataset_config=None, add_post_process_config=None)
    stub("glass.postprocess", build_post_processor=None)
    stub("glass.modeling.recognition.text_encoder", TextEncoder=None)

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m
    load("glass.utils.common_utils", "glass/utils/common_utils.py")
    return load("ref_glass_runner", "glass/inference/glass_runner.py"), structures


def main():
    from golden_common import make_runner_case
    gr, st = load_reference()
    R = gr.GlassRunner
    ns = types.SimpleNamespace(min_target_size=1200, max_target_size=1600, max_upscale_ratio=2)
    ns.get_inference_scale_ratio = lambda shape: R.get_inference_scale_ratio(ns, shape)
    shapes = [(1024, 1024, 3), (300, 500, 3), (1200, 900, 3), (1600, 1601, 3), (4000, 3000, 3), (599, 600, 3), (600, 601, 3),
              (1199, 100, 3), (1400, 1400, 3), (97, 130, 3)]
    ratios = [float(R.get_inference_scale_ratio(ns, s)) for s in shapes]
    cases = []
    for seed, hw, fmt in [(0, (97, 130), "BGR"), (1, (150, 111), "RGB"), (2, (64, 80), "GREY"), (3, (1300, 40), "BGR")]:
        image, boxes, scores = make_runner_case(seed, hw)
        seen = {}

        class Model:
            device = torch.device("cpu")

            def __call__(self, inputs):
                seen["image"], seen["height"], seen["width"] = inputs[0]["image"], inputs[0]["height"], inputs[0]["width"]
                sc = seen["image"].shape[1] / hw[0]
                inst = st.Instances((seen["height"], seen["width"]), pred_boxes=st.RotatedBoxes(boxes.clone() * torch.tensor([sc, sc, sc, sc, 1.0])),
                                    scores=scores.clone())
                return [{"instances": inst}]

        def post(preds):
            seen["post_in_size"], seen["post_in_boxes"] = tuple(preds.image_size), preds.pred_boxes.tensor.clone()
            return preds[preds.scores > 0.5]

        run = types.SimpleNamespace(min_target_size=ns.min_target_size, max_target_size=ns.max_target_size,
                                    max_upscale_ratio=ns.max_upscale_ratio, input_format=fmt, model=Model(),
                                    post_processor=post, logger=logging.getLogger("golden"))
        run.get_inference_scale_ratio = lambda shape: R.get_inference_scale_ratio(run, shape)
        run._image_to_tensor = lambda img, dev, interpolation="bilinear": R._image_to_tensor(run, img, dev, interpolation)
        try:
            out = R.__call__(run, image)
        except ValueError as e:   # "RGB": image[:, :, ::-1] has a negative stride, torch.as_tensor refuses it (:84, :133)
            cases.append({"seed": seed, "hw": hw, "format": fmt, "raises": str(e)[:60]})
            print(f"case {seed} {fmt}: the reference raises ValueError ({str(e)[:50]}...)")
            continue
        t = seen["image"]
        cases.append({"seed": seed, "hw": hw, "format": fmt, "model_hw": (seen["height"], seen["width"]),
                      "tensor_shape": tuple(t.shape), "tensor_sum": float(t.double().sum()),
                      "tensor_sample": t[:, ::7, ::5].contiguous().clone(),
                      "post_in_size": seen["post_in_size"], "post_in_boxes": seen["post_in_boxes"],
                      "out_size": tuple(out.image_size), "out_boxes": out.pred_boxes.tensor.clone(), "out_scores": out.scores.clone()})
        print(f"case {seed} {fmt} {hw}: model saw {tuple(t.shape)}, {len(out)} of {len(boxes)} instances returned")
    torch.save({"shapes": shapes, "ratios": ratios, "cases": cases}, os.path.join(ROOT, "tests", "golden", "runner.pt"))


if __name__ == "__main__":
    main()
