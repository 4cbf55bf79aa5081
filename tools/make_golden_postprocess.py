"""Generate tests/golden/postprocess.pt by running the REFERENCE's own post-processor classes
(glass/postprocess/post_processor_rotated_boxes.py + post_processor_academic.py's text filter) on seeded detections.

Authoring container only (needs /root/reference).  detectron2 is not installable offline, so the reference files are
imported under stubs: ``Instances`` is a ~30-line field bag, ``nms_rotated`` / ``pairwise_iou_rotated`` are the
oracle's C restatements of the detectron2 operators (pinned by detectron2's upstream KATs, tests/test_oracle_d2_ops.py).
Everything GLASS-specific -- the merge loop, the pair masks, the cv2.minAreaRect re-orientation, the write-back order,
the thresholds -- is the reference's code, unmodified.

    python tools/make_golden_postprocess.py
"""
import importlib.util
import os
import sys
import types

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


class Boxes5:
    def __init__(self, t):
        self.tensor = t

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, item):
        return Boxes5(self.tensor[item].reshape(-1, 5))


class Instances:
    def __init__(self, image_size, **fields):
        self._image_size = image_size
        self._fields = dict(fields)

    def __getattr__(self, name):
        if name.startswith("_") or name not in self._fields:
            raise AttributeError(name)
        return self._fields[name]

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self._fields[name] = val

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        return 0

    def __getitem__(self, item):
        if isinstance(item, torch.Tensor) and item.dtype == torch.bool:
            item = torch.nonzero(item).squeeze(1)
        return Instances(self._image_size, **{k: v[item] for k, v in self._fields.items()})


def install_stubs():
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

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("detectron2")
    mod("detectron2.layers")
    mod("detectron2.layers.nms", nms_rotated=lambda boxes, scores, iou_threshold: d2_ops.nms_rotated(boxes, scores, iou_threshold))
    mod("detectron2.structures")
    mod("detectron2.structures.instances", Instances=Instances)
    mod("detectron2.structures.boxes", BoxMode=object, Boxes=object, pairwise_ioa=None, pairwise_intersection=None)
    mod("detectron2.structures.rotated_boxes", pairwise_iou_rotated=lambda a, b: d2_ops.box_iou_rotated(a, b))
    mod("detectron2.utils")
    mod("detectron2.utils.registry", Registry=Registry)


def load_ref_pkg():
    """Import glass.structures.boxes and glass.postprocess.post_processor_rotated_boxes as a package (relative imports)."""
    for name in ("glass", "glass.structures", "glass.postprocess"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m
    load("glass.structures.boxes", "glass/structures/boxes.py")
    pp = load("glass.postprocess.post_processor_rotated_boxes", "glass/postprocess/post_processor_rotated_boxes.py")
    te = load("ref_text_encoder", "glass/modeling/recognition/text_encoder.py")
    return pp, te


def make_cfg():
    """cfg.POST_PROCESSING defaults (glass/config.py:176-214) + the keys the constructor reads."""
    pp = types.SimpleNamespace(SKIP_ALL=False, MIN_BOX_DIMENSION=2, MERGE_IOA_THRESH=0.3, PAIRS_HEIGHT_RATIO_THRESH=0.35,
                               BOX_PX_PADDING=[0, 0, 0, 0], VALID_CONFIDENCE=0.15, DETECT_THRESHOLD=0.25,
                               TEXT_THRESHOLD=0.25, MAX_ANGLE_DIFF=15)
    return types.SimpleNamespace(POST_PROCESSING=pp,
                                 MODEL=types.SimpleNamespace(ROI_HEADS=types.SimpleNamespace(CLASS_NAMES=["word"])),
                                 INPUT=types.SimpleNamespace(MAX_SIZE_TEST=1600))


from tests.golden_common import make_postprocess_case as make_case  # noqa: E402  (shared with the GPU tests)


def main():
    install_stubs()
    pp_mod, te_mod = load_ref_pkg()
    cfg = make_cfg()
    post = pp_mod.PostProcessorRotatedBoxes(cfg)
    # the reference's TextEncoder, for the academic text-score filter (post_processor_academic.py:31-32 via
    # text_evaluator.get_instances_text:323-331 = max over classes, then decode_prod_v2's word score)
    y = yaml.safe_load(open(os.path.join(REF, "configs/glass_pretrain.yaml")))
    rh = y["MODEL"]["ROI_RECOGNIZER_HEAD"]
    head = types.SimpleNamespace(NAME="RecognizerRCNNHeadV3", MAX_WORD_LENGTH=rh["MAX_WORD_LENGTH"], CHARACTER_SET=rh["CHARACTER_SET"],
                                 UNK_SYMBOL_PRED=rh.get("UNK_SYMBOL_PRED", False), LABELS_TYPE="attention", IGNORE_TEXT=[],
                                 IGNORE_EMPTY_TEXT=True)
    enc = te_mod.TextEncoder(types.SimpleNamespace(MODEL=types.SimpleNamespace(ROI_RECOGNIZER_HEAD=head)))

    cases = []
    specs = [(0, 6, 10), (1, 12, 20), (2, 20, 30), (3, 1, 0), (4, 0, 5), (5, 25, 10), (6, 3, 60), (7, 16, 36)]
    for seed, n_lines, n_iso in specs:
        boxes, scores = make_case(seed, n_lines, n_iso)
        boxes, scores = boxes[:100], scores[:100]  # DETECTIONS_PER_IMAGE
        n = len(boxes)
        g = torch.Generator().manual_seed(1000 + seed)
        # per-step class probabilities: a peaked distribution per step, stop symbol (class 1) somewhere
        logits = torch.randn(n, 26, 97, generator=g) * 2.5
        probs = torch.softmax(logits * 5.0, dim=2)
        for i in range(n):
            stop = int(torch.randint(1, 12, (1,), generator=g))
            probs[i, stop] = 0.002
            probs[i, stop, 1] = 0.808
        pmax, pidx = probs.max(dim=2)
        dec = enc.decode_prod_v2(pred_probs=pmax.numpy().copy(), pred_indices=pidx.numpy().copy())
        text_scores = torch.tensor([float(d["score"]) for d in dec], dtype=torch.float32)

        inst = Instances((1024, 1024), pred_boxes=Boxes5(boxes.clone()), scores=scores.clone(), orig_idx=torch.arange(n),
                         text_scores=text_scores.clone())
        out = post(inst)  # PostProcessorRotatedBoxes.__call__
        rb_idx, rb_boxes, rb_poly = out.orig_idx.clone(), out.pred_boxes.tensor.clone(), out.pred_polygons.clone()
        out = out[out.text_scores >= cfg.POST_PROCESSING.TEXT_THRESHOLD]  # PostProcessorAcademic.__call__ :31-32
        cases.append({"seed": seed, "boxes": boxes, "scores": scores, "text_prob_max": pmax, "text_prob_idx": pidx,
                      "text_scores": text_scores,
                      "rb_idx": rb_idx, "rb_boxes": rb_boxes, "rb_polygons": rb_poly,
                      "idx": out.orig_idx.clone(), "out_boxes": out.pred_boxes.tensor.clone(),
                      "out_polygons": out.pred_polygons.clone()})
        print(f"case {seed}: {n} detections -> {len(rb_idx)} after merge/filters -> {len(out)} after the text filter")
    # the static helpers, on their own
    b, _ = make_case(99, 10, 10)
    poly = pp_mod.PostProcessorRotatedBoxes.boxes_to_polygons(b)
    back = pp_mod.PostProcessorRotatedBoxes.polygons_to_rotated_boxes(poly, orientations=b[:, 4])
    torch.save({"cases": cases, "helper_boxes": b, "helper_polygons": poly, "helper_roundtrip": back,
                "cv2": __import__("cv2").__version__},
               os.path.join(ROOT, "tests", "golden", "postprocess.pt"))
    print("saved", len(cases), "cases")


if __name__ == "__main__":
    main()
