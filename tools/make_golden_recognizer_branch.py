"""Generate tests/golden/recognizer_branch.pt with the REFERENCE's own recognizer branch, end to end:
``MaskRotatedRecognizerHybridHead._forward_recognizer`` (glass/modeling/fusion/recognizers_hybrid_head.py:513-569) driving
the reference's own ``P2P3Fusion``, ``ResNetFeatureExtractor``, ``MultiAspectGCAttention`` and a ``RecognizerRCNNHeadV3``
(glass/modeling/recognition/recognizer_head_v2.py:291-345, inference :150-163) that the reference's own builders
assemble from the REFERENCE's configs/glass_pretrain.yaml (``build_recognizer_backbonev2`` -> CNN_V1_1,
``build_recognizer_encoderv2`` -> BiLSTMBlockV2, ``build_recognizer_decoderv2`` -> ASTER_V2).

This pins the WIRING of rows a9-a16: which map the recognizer pooler reads, the image pooler's output
size, the [local | global] concatenation order, the early-break decoding over all words of the image, the split of the
predictions back onto the instances -- each stage on its own is already pinned by tests/golden/{hybrid_net,fusion_net,
p2p3,cnn_v1_1,encoder,decoder}.pt.

Authoring container only.  detectron2 is not installable offline; stubbed: ``configurable`` (calls ``from_config`` when
the first argument is a config node, like detectron2's), ``Registry``, ``Conv2d`` / ``get_norm`` (published semantics),
``ROIPooler`` (the oracle's restatement, oracle/d2_ops.py, pinned by detectron2's upstream known-answer tests),
``Instances`` / ``RotatedBoxes`` (this repo's field bags).  The orchestration is the reference's code, unmodified.

    python tools/make_golden_recognizer_branch.py
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


class Conv2d(nn.Conv2d):
    """detectron2.layers.wrappers.Conv2d: conv -> norm -> activation."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


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
        """detectron2.config.configurable for __init__: ``Cls(cfg, ...)`` goes through ``Cls.from_config``."""
        def wrapped(self, *args, **kwargs):
            if args and isinstance(args[0], CfgNode):
                init_func(self, **type(self).from_config(*args, **kwargs))
            else:
                init_func(self, *args, **kwargs)
        return wrapped

    class ROIPooler:
        """detectron2.modeling.poolers.ROIPooler -> the oracle's restatement."""

        def __init__(self, output_size, scales, sampling_ratio, pooler_type):
            assert pooler_type == "ROIAlignRotated"
            self.output_size, self.scales, self.sampling_ratio = tuple(output_size), list(scales), sampling_ratio

        def __call__(self, x, box_lists):
            return d2_ops.roi_pooler(x, [b.tensor for b in box_lists], self.output_size, self.scales, self.sampling_ratio)

    class ShapeSpec:
        def __init__(self, channels=None, height=None, width=None, stride=None):
            self.channels, self.height, self.width, self.stride = channels, height, width, stride

    class StandardROIHeads(nn.Module):
        pass

    wi = stub("fvcore.nn.weight_init", c2_msra_fill=lambda m: None, c2_xavier_fill=lambda m: None)
    stub("fvcore")
    stub("fvcore.nn", weight_init=wi)
    stub("detectron2")
    stub("detectron2.config", configurable=configurable)
    stub("detectron2.layers", Conv2d=Conv2d, ShapeSpec=ShapeSpec, get_norm=lambda norm, c: nn.BatchNorm2d(c) if norm else None)
    stub("detectron2.modeling")
    stub("detectron2.modeling.poolers", ROIPooler=ROIPooler)
    stub("detectron2.modeling.roi_heads")
    stub("detectron2.modeling.roi_heads.box_head", build_box_head=None)
    stub("detectron2.modeling.roi_heads.mask_head", build_mask_head=None, ROI_MASK_HEAD_REGISTRY=Registry("ROI_MASK_HEAD"))
    stub("detectron2.modeling.roi_heads.roi_heads", ROI_HEADS_REGISTRY=Registry("ROI_HEADS"), StandardROIHeads=StandardROIHeads,
         select_foreground_proposals=None)
    stub("detectron2.modeling.roi_heads.rotated_fast_rcnn", RotatedFastRCNNOutputLayers=None)
    stub("detectron2.structures", ImageList=structures.ImageList, Instances=structures.Instances,
         RotatedBoxes=structures.RotatedBoxes, pairwise_iou_rotated=None, Boxes=object)
    stub("detectron2.utils")
    stub("detectron2.utils.comm", is_main_process=lambda: True)
    stub("detectron2.utils.events", get_event_storage=None)
    stub("detectron2.utils.registry", Registry=Registry)
    for name in ("glass", "glass.modeling", "glass.modeling.fusion", "glass.modeling.recognition", "glass.modeling.roi_heads",
                 "glass.structures"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    # glass modules the hybrid head imports but the recognizer branch never reaches
    stub("glass.modeling.recognition.recognizer_pooler_pad", build_recognizer_pooler_pad=None)
    stub("glass.modeling.roi_heads.rotated_fast_rcnn", RotatedFastRCNNOutputLayers=None, overwrite_orientations_on_boxes=None)
    stub("glass.modeling.roi_heads.rotated_head", add_ground_truth_to_proposals=None)
    stub("glass.structures.boxes", box_to_rbox=None, rbox_to_box=None)

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m
    r = "glass/modeling/recognition/"
    for n in ("text_encoder", "recognizer_backbone", "recognizer_encoder", "prediction_aster", "recognizer_decoder",
              "recognizer_head_v2"):
        load("glass.modeling.recognition." + n, r + n + ".py")
    sys.modules["glass.modeling.recognition.prediction_aster"].device = torch.device("cpu")   # module-global (:11)
    load("glass.modeling.fusion.fusion_modules", "glass/modeling/fusion/fusion_modules.py")
    load("glass.modeling.fusion.local_feature_extraction", "glass/modeling/fusion/local_feature_extraction.py")
    hh = load("glass.modeling.fusion.recognizers_hybrid_head", "glass/modeling/fusion/recognizers_hybrid_head.py")
    return hh, ShapeSpec, ROIPooler, structures


def reference_config():
    """The reference's own pretrain config over this repo's defaults, plus the keys only its constructors read."""
    from glass_text_spotting_b200 import config
    cfg = config.load_config(os.path.join(REF, "configs/glass_pretrain.yaml"))
    cfg.merge({"VIS_PERIOD": 0, "MODEL": {"ROI_RECOGNIZER_HEAD": {
        "IGNORE_EMPTY_TEXT": True, "IGNORE_TEXT": ["###"], "MAX_BATCH_SIZE": 1, "LOSS_WEIGHT": 2.0,
        "SAMPLE_WORDS_STRATEGY": "random", "SAMPLE_WORDS_STRATEGY_PROB": 0.3}}})
    return cfg


def build_reference_branch(seed: int):
    """-> (namespace standing in for the ROI head's ``self``, hybrid-head module, structures)."""
    from golden_common import seeded_fill
    hh, ShapeSpec, ROIPooler, structures = load_reference()
    cfg = reference_config()
    fus = sys.modules["glass.modeling.fusion.fusion_modules"]
    lfe = sys.modules["glass.modeling.fusion.local_feature_extraction"]
    head_mod = sys.modules["glass.modeling.recognition.recognizer_head_v2"]
    rec = cfg.MODEL.ROI_RECOGNIZER_HEAD
    # recognizers_hybrid_head.py:444-511 (_init_recognizer_head), with the values its cfg reads resolve to
    ns = types.SimpleNamespace(
        training=False,
        recognizer_in_features=list(rec.IN_FEATURES),
        recognizer_pooler=ROIPooler(output_size=[rec.POOLER_RESOLUTION_HEIGHT, rec.POOLER_RESOLUTION_WIDTH],
                                    scales=(1.0 / 4,), sampling_ratio=rec.POOLER_SAMPLING_RATIO, pooler_type=rec.POOLER_TYPE),
        img_pooler=ROIPooler(output_size=[rec.POOLER_RESOLUTION_HEIGHT * 16, rec.POOLER_RESOLUTION_WIDTH * 4], scales=[1],
                             sampling_ratio=cfg.MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO,
                             pooler_type=cfg.MODEL.ROI_BOX_HEAD.POOLER_TYPE),
        recognizer_feature_fusion=fus.P2P3Fusion(256).eval(),
        hybrid_net=lfe.ResNetFeatureExtractor(3, cfg.MODEL.LOCAL_FEATURE_EXTRACTOR.NUM_FEATURES).eval(),
        fusion_net=fus.MultiAspectGCAttention(inplanes=512, ratio=cfg.MODEL.HYBRID_FUSION.RATIO,
                                              headers=cfg.MODEL.HYBRID_FUSION.HEADERS, outplane=256,
                                              fusion_type=cfg.MODEL.HYBRID_FUSION.FUSION_TYPE).eval(),
        recognizer_head=head_mod.RecognizerRCNNHeadV3(cfg, ShapeSpec(channels=256, height=8, width=32)).eval(),
    )
    for i, name in enumerate(("recognizer_feature_fusion", "hybrid_net", "fusion_net", "recognizer_head")):
        seeded_fill(getattr(ns, name), seed + i)
    return ns, hh, structures


EOS_EXTRA = 2.2   # on top of force_eos_bias: the four words then emit class 0 at steps 2, 2, 9, 2 -> break after 10 steps


def main():
    from golden_common import force_eos_bias, make_recognizer_branch_inputs
    cases = []
    with torch.no_grad():
        for seed, k, eos in [(0, 3, False), (1, 1, False), (2, 0, False), (3, 4, True)]:
            ns, hh, st = build_reference_branch(500 + 10 * seed)
            if eos:   # make the early break of prediction_aster.py:93-95 fire inside the branch
                force_eos_bias(ns.recognizer_head.decoder.recognizer)
                ns.recognizer_head.decoder.recognizer.decoder.fc.bias[0] += EOS_EXTRA
            image, p2, p3, boxes = make_recognizer_branch_inputs(seed, k)
            images = st.ImageList(image[None], [tuple(image.shape[-2:])])
            inst = st.Instances(tuple(image.shape[-2:]), pred_boxes=st.RotatedBoxes(boxes.clone()),
                                pred_classes=torch.zeros(k, dtype=torch.int64))
            out = hh.MaskRotatedRecognizerHybridHead._forward_recognizer(ns, images, {"p2": p2, "p3": p3}, [inst])
            has = out[0].has("pred_text_prob")
            probs = out[0].pred_text_prob.clone() if has else None
            keys = {n: sorted(getattr(ns, n).state_dict().keys())
                    for n in ("recognizer_feature_fusion", "hybrid_net", "fusion_net", "recognizer_head")}
            cases.append({"seed": seed, "k": k, "eos": eos, "has_text": has, "pred_text_prob": probs, "state_keys": keys})
            if has:
                steps = int((probs.sum(2) > 0).sum(1).max())
                print(f"case {seed}: {k} words -> pred_text_prob {tuple(probs.shape)}, {steps} decoding steps before the break, "
                      f"argmax[0] = {probs[0].argmax(1)[:8].tolist()}")
            else:
                print(f"case {seed}: {k} words -> no pred_text_prob field (recognizer_head_v2.py:151-152)")
    torch.save({"cases": cases}, os.path.join(ROOT, "tests", "golden", "recognizer_branch.pt"))


if __name__ == "__main__":
    main()
