"""Reference configuration files -> B200 modules.

The reference builds everything from a yacs ``cfg`` (``cls(cfg, input_shape)`` / ``@configurable from_config``):
``get_cfg()`` + ``add_e2e_config`` / ``add_glass_config`` / ``add_post_process_config`` (glass/config.py) merged with
one of ``configs/*.yaml``, then ``build_model(cfg)`` and ``build_post_processor(cfg)``
(glass/inference/glass_runner.py:52-70).  yacs / detectron2 are not installable offline, so this module reads the same
YAML files into a plain attribute tree over the same defaults (only the keys the inference path reads) and maps them
onto the constructors of ``B200GlassRCNN`` / ``B200PostProcessor`` / ``B200GlassRunner``.

The B200 path implements ONE architecture -- the one all four shipped configs select.  A config that asks for anything
else (another backbone, head, recognizer part, norm-free FPN, ...) is refused with the offending key, never silently
approximated.
"""
import copy
import os
from typing import Any, Dict, Mapping, Optional, Union

import yaml

# Defaults of the keys the inference path reads: detectron2 v0.6 config/defaults.py for the stock keys, glass/config.py
# (line numbers in the comments) for the GLASS ones.
_CHARSET = '0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!"#$%&\'()*+,-./:;<=>?@[\\]^_`{|}~ '
DEFAULTS: Dict[str, Any] = {
    "MODEL": {
        "META_ARCHITECTURE": "GeneralizedRCNN",
        "PIXEL_MEAN": [103.530, 116.280, 123.675],
        "PIXEL_STD": [1.0, 1.0, 1.0],
        "MASK_ON": False,
        "ROTATED_BOXES_ON": False,                      # :25
        "ORIENTATION_ON": False,                        # :26
        "RECOGNIZER_ON": False,                         # :91
        "BACKBONE": {"NAME": "build_resnet_backbone", "FREEZE_AT": 2},
        "RESNETS": {"DEPTH": 50, "OUT_FEATURES": ["res4"], "NUM_GROUPS": 1, "NORM": "FrozenBN", "WIDTH_PER_GROUP": 64,
                    "STRIDE_IN_1X1": True, "RES5_DILATION": 1, "RES2_OUT_CHANNELS": 256, "STEM_OUT_CHANNELS": 64,
                    "DEFORM_ON_PER_STAGE": [False, False, False, False]},
        "FPN": {"IN_FEATURES": [], "OUT_CHANNELS": 256, "NORM": "", "FUSE_TYPE": "sum"},
        "PROPOSAL_GENERATOR": {"NAME": "RPN", "MIN_SIZE": 0},
        "ANCHOR_GENERATOR": {"NAME": "DefaultAnchorGenerator", "SIZES": [[32, 64, 128, 256, 512]],
                             "ASPECT_RATIOS": [[0.5, 1.0, 2.0]], "ANGLES": [[-90, 0, 90]], "OFFSET": 0.0},
        "RPN": {"HEAD_NAME": "StandardRPNHead", "IN_FEATURES": ["res4"], "BBOX_REG_WEIGHTS": [1.0, 1.0, 1.0, 1.0],
                "PRE_NMS_TOPK_TEST": 6000, "POST_NMS_TOPK_TEST": 1000, "NMS_THRESH": 0.7},
        "ROI_HEADS": {"NAME": "Res5ROIHeads", "NUM_CLASSES": 80, "IN_FEATURES": ["res4"], "SCORE_THRESH_TEST": 0.05,
                      "NMS_THRESH_TEST": 0.5, "CLASS_NAMES": ["word"]},                                     # :53
        "ROI_BOX_HEAD": {"NAME": "", "BBOX_REG_WEIGHTS": [10.0, 10.0, 5.0, 5.0], "POOLER_RESOLUTION": 14,
                         "POOLER_SAMPLING_RATIO": 0, "POOLER_TYPE": "ROIAlignV2", "NUM_FC": 0, "FC_DIM": 1024,
                         "NUM_CONV": 0, "CONV_DIM": 256, "NORM": "", "CLS_AGNOSTIC_BBOX_REG": False},
        "ROI_MASK_HEAD": {"NAME": "MaskRCNNConvUpsampleHead", "POOLER_RESOLUTION": 14, "POOLER_SAMPLING_RATIO": 0,
                          "NUM_CONV": 0, "CONV_DIM": 256, "NORM": "", "CLS_AGNOSTIC_MASK": False,
                          "POOLER_TYPE": "ROIAlignV2", "MASK_INFERENCE": False,                              # :170
                          "IN_FEATURES": ["p2", "p3", "p4", "p5", "p6"]},                                  # :99
        "ROI_RECOGNIZER_HEAD": {                                                                              # :126-168
            "NAME": "", "LABELS_TYPE": "attention", "MAX_WORD_LENGTH": 50, "CHARACTER_SET": _CHARSET,
            "UNK_SYMBOL_PRED": False, "POOLER_RESOLUTION_WIDTH": 32, "POOLER_RESOLUTION_HEIGHT": 32,
            "IN_FEATURES": ["p2", "p3", "p4", "p5", "p6"], "CLASS_IND": 0, "POOLER_TYPE": "ROIAlignRotated", "NORM": "BN",
            "POOLER_SAMPLING_RATIO": 0, "SAMPLING_RATIO": 0, "CONV_DIM": 256, "SENSITIVE": True,
            "RECOGNIZER_HEAD": {"BACKBONE": {"NAME": "CNN_V1_2"},
                                "ENCODER": {"NAME": "BiLSTMBlockV2", "NUM_OF_LAYERS": 2, "HEIGHT_REDUCTION": "mean"},
                                "DECODER": {"NAME": "ASTER_V2"}}},
        "LOCAL_FEATURE_EXTRACTOR": {"NAME": "ResNet_FeatureExtractor", "NUM_FEATURES": 256},                # :40-42
        "HYBRID_FUSION": {"NAME": "MultiAspectGCAttention", "NUM_FEATURES": 256, "RATIO": 0.5, "HEADERS": 8,
                          "FUSION_TYPE": "channel_add"},                                                      # :44-51
        "ROI_ORIENTATION_HEAD": {"APPLY_TO_BOXES": False},                                                   # :56
    },
    "INPUT": {"FORMAT": "BGR", "MIN_SIZE_TEST": 1600, "MAX_SIZE_TEST": 1600, "MAX_UPSCALE_RATIO": 2},        # :63-67
    "TEST": {"DETECTIONS_PER_IMAGE": 100},
    "POST_PROCESSING": {                                                                                      # :176-214
        "NAME": "PostProcessorAcademic", "SKIP_ALL": False, "BOX_INFLATE_RATIO": 0.05, "BOX_PX_PADDING": [0, 0, 0, 0],
        "MIN_BOX_DIMENSION": 2, "MERGE_IOA_THRESH": 0.3, "PAIRS_HEIGHT_RATIO_THRESH": 0.35, "LOW_CONFIDENCE": 0.01,
        "VALID_CONFIDENCE": 0.15, "DETECT_THRESHOLD": 0.25, "TEXT_THRESHOLD": 0.25, "MAX_ANGLE_DIFF": 15},
}


class CfgNode(dict):
    """Attribute access over nested dicts (``cfg.MODEL.RPN.NMS_THRESH``), ``hasattr``-friendly like yacs' CfgNode --
    the reference probes optional keys with hasattr (glass_rcnn.py:43-53)."""

    def __init__(self, d: Optional[Mapping] = None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = CfgNode(v) if isinstance(v, Mapping) else copy.deepcopy(v)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def merge(self, other: Mapping) -> "CfgNode":
        for k, v in other.items():
            if isinstance(v, Mapping) and isinstance(self.get(k), Mapping):
                self[k].merge(v)
            else:
                self[k] = CfgNode(v) if isinstance(v, Mapping) else copy.deepcopy(v)
        return self


def get_cfg() -> CfgNode:
    """The defaults the reference starts from (detectron2 ``get_cfg`` + glass/config.py's ``add_*_config``)."""
    return CfgNode(DEFAULTS)


def load_config(source: Union[str, os.PathLike, Mapping]) -> CfgNode:
    """``cfg.merge_from_file(path)`` over the defaults; also accepts an already-parsed mapping.  Keys the inference path
    does not read (SOLVER, DATASETS, ...) are carried along untouched."""
    if isinstance(source, Mapping):
        loaded = source
    else:
        with open(source) as fp:
            loaded = yaml.safe_load(fp)
    if "_BASE_" in loaded:
        raise NotImplementedError("_BASE_ inheritance is not used by the reference's configs and is not supported")
    return get_cfg().merge(loaded)


class UnsupportedConfig(ValueError):
    pass


def _require(cfg: CfgNode, dotted: str, expected) -> None:
    node: Any = cfg
    for part in dotted.split("."):
        node = node[part]
    ok = node in expected if isinstance(expected, (set, frozenset)) else node == expected
    if not ok:
        raise UnsupportedConfig(f"{dotted} = {node!r}: the B200 path implements {expected!r} "
                                f"(the architecture of the reference's shipped configs) and has no other variant")


def check_supported(cfg: CfgNode) -> None:
    """Every architectural choice must be the one the kernels implement; see the module docstring."""
    m = "MODEL."
    _require(cfg, m + "META_ARCHITECTURE", {"GeneralizedRCNN", "GlassRCNN"})
    _require(cfg, m + "BACKBONE.NAME", "build_resnet_fpn_backbone")
    _require(cfg, m + "RESNETS.DEPTH", 50)
    _require(cfg, m + "RESNETS.OUT_FEATURES", ["res2", "res3", "res4", "res5"])
    _require(cfg, m + "RESNETS.NUM_GROUPS", 1)
    _require(cfg, m + "RESNETS.STRIDE_IN_1X1", True)
    _require(cfg, m + "RESNETS.RES5_DILATION", 1)
    _require(cfg, m + "RESNETS.RES2_OUT_CHANNELS", 256)
    _require(cfg, m + "RESNETS.STEM_OUT_CHANNELS", 64)
    _require(cfg, m + "RESNETS.NORM", {"SyncBN", "BN", "FrozenBN"})      # identical in eval (running statistics)
    _require(cfg, m + "RESNETS.DEFORM_ON_PER_STAGE", [False, False, False, False])
    _require(cfg, m + "FPN.IN_FEATURES", ["res2", "res3", "res4", "res5"])
    _require(cfg, m + "FPN.OUT_CHANNELS", 256)
    _require(cfg, m + "FPN.NORM", {"SyncBN", "BN", "FrozenBN"})          # conv bias off + norm
    _require(cfg, m + "FPN.FUSE_TYPE", "sum")
    _require(cfg, m + "PROPOSAL_GENERATOR.NAME", "RotatedRPN")
    _require(cfg, m + "ANCHOR_GENERATOR.NAME", "RotatedAnchorGenerator")
    _require(cfg, m + "RPN.HEAD_NAME", "StandardRPNHead")
    _require(cfg, m + "RPN.IN_FEATURES", ["p2", "p3", "p4", "p5", "p6"])
    _require(cfg, m + "ROI_HEADS.NAME", "MaskRotatedRecognizerHybridHead")
    _require(cfg, m + "ROI_HEADS.IN_FEATURES", ["p2", "p3", "p4", "p5", "p6"])
    _require(cfg, m + "ROI_HEADS.NUM_CLASSES", 1)
    _require(cfg, m + "ROTATED_BOXES_ON", True)
    _require(cfg, m + "ROI_BOX_HEAD.NAME", "FastRCNNConvFCHead")
    _require(cfg, m + "ROI_BOX_HEAD.NUM_FC", 2)
    _require(cfg, m + "ROI_BOX_HEAD.NUM_CONV", 0)
    _require(cfg, m + "ROI_BOX_HEAD.POOLER_TYPE", "ROIAlignRotated")
    _require(cfg, m + "ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG", False)
    r = m + "ROI_RECOGNIZER_HEAD."
    _require(cfg, r + "NAME", "RecognizerRCNNHeadV3")
    _require(cfg, r + "LABELS_TYPE", "attention")
    _require(cfg, r + "IN_FEATURES", ["p2", "p3"])
    _require(cfg, r + "POOLER_TYPE", "ROIAlignRotated")
    _require(cfg, r + "UNK_SYMBOL_PRED", False)
    _require(cfg, r + "CONV_DIM", 256)
    _require(cfg, r + "RECOGNIZER_HEAD.BACKBONE.NAME", "CNN_V1_1")
    _require(cfg, r + "RECOGNIZER_HEAD.ENCODER.NAME", "BiLSTMBlockV2")
    _require(cfg, r + "RECOGNIZER_HEAD.ENCODER.NUM_OF_LAYERS", 2)
    _require(cfg, r + "RECOGNIZER_HEAD.ENCODER.HEIGHT_REDUCTION", "mean")
    _require(cfg, r + "RECOGNIZER_HEAD.DECODER.NAME", "ASTER_V2")
    _require(cfg, m + "LOCAL_FEATURE_EXTRACTOR.NAME", "ResNetFeatureExtractor")
    _require(cfg, m + "LOCAL_FEATURE_EXTRACTOR.NUM_FEATURES", 256)
    _require(cfg, m + "HYBRID_FUSION.NAME", "MultiAspectGCAttention")
    _require(cfg, m + "HYBRID_FUSION.NUM_FEATURES", 256)
    _require(cfg, m + "HYBRID_FUSION.RATIO", 0.5)
    _require(cfg, m + "HYBRID_FUSION.HEADERS", 8)
    _require(cfg, m + "HYBRID_FUSION.FUSION_TYPE", "channel_add")
    if cfg.MODEL.ROI_MASK_HEAD.MASK_INFERENCE:
        _require(cfg, m + "ROI_MASK_HEAD.NAME", "RotatedMaskRCNNConvUpsampleHead")
        _require(cfg, m + "ROI_MASK_HEAD.NUM_CONV", 4)
        _require(cfg, m + "ROI_MASK_HEAD.POOLER_RESOLUTION", 14)
        _require(cfg, m + "ROI_MASK_HEAD.POOLER_TYPE", "ROIAlignRotated")
        _require(cfg, m + "ROI_MASK_HEAD.NORM", "")
    # hard numeric limits of the kernels (a YAML that omits these keys inherits detectron2's defaults, e.g. RPN top-k
    # 6000 / 1000, which the device path cannot hold): refuse here, by key, not at the first image
    M = cfg.MODEL
    rec = M.ROI_RECOGNIZER_HEAD
    n_anchor_levels = len(M.ANCHOR_GENERATOR.SIZES)
    for dotted, value, limit, why in (
            ("MODEL.RPN.PRE_NMS_TOPK_TEST", int(M.RPN.PRE_NMS_TOPK_TEST), 1024, "per-level top-k is sorted in one CTA (csrc/detect.cu)"),
            ("MODEL.RPN.PRE_NMS_TOPK_TEST x levels", n_anchor_levels * int(M.RPN.PRE_NMS_TOPK_TEST), 8192,
             "the proposal NMS holds at most 8192 candidates per image"),
            ("MODEL.RPN.POST_NMS_TOPK_TEST", int(M.RPN.POST_NMS_TOPK_TEST), 128, "NMS max_keep <= 128"),
            ("TEST.DETECTIONS_PER_IMAGE", int(cfg.TEST.DETECTIONS_PER_IMAGE), 128,
             "NMS max_keep and the device post-processor hold <= 128 detections per image"),
            ("MODEL.ROI_RECOGNIZER_HEAD.POOLER_RESOLUTION_WIDTH", int(rec.POOLER_RESOLUTION_WIDTH), 32,
             "the decoder attends over T <= 32 positions"),
            ("len(MODEL.ROI_RECOGNIZER_HEAD.CHARACTER_SET) + 2", len(rec.CHARACTER_SET) + 2, 128, "decoder classes <= 128"),
            ("MODEL.ROI_RECOGNIZER_HEAD.MAX_WORD_LENGTH + 1", int(rec.MAX_WORD_LENGTH) + 1, 64, "decoding steps <= 64")):
        if value > limit or value < 1:
            raise UnsupportedConfig(f"{dotted} = {value}: outside the device path's range [1, {limit}] ({why})")
    if int(rec.POOLER_RESOLUTION_HEIGHT) != 8 or int(rec.POOLER_RESOLUTION_WIDTH) != 32:
        raise UnsupportedConfig("MODEL.ROI_RECOGNIZER_HEAD.POOLER_RESOLUTION_HEIGHT/WIDTH = "
                                f"{rec.POOLER_RESOLUTION_HEIGHT}/{rec.POOLER_RESOLUTION_WIDTH}: the recognizer is built for the "
                                "8 x 32 pooler of the shipped configs (local crops 128 x 128, T = 32)")
    if len(cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS) != 1 or len(cfg.MODEL.ANCHOR_GENERATOR.ANGLES) != 1:
        raise UnsupportedConfig("MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS / ANGLES: one list shared by all levels expected")
    if len(cfg.MODEL.ANCHOR_GENERATOR.SIZES) != 5 or any(len(s) != 1 for s in cfg.MODEL.ANCHOR_GENERATOR.SIZES):
        raise UnsupportedConfig("MODEL.ANCHOR_GENERATOR.SIZES: one size per level for p2..p6 expected")


def model_kwargs(cfg: CfgNode) -> Dict[str, Any]:
    """cfg -> keyword arguments of ``B200GlassRCNN`` (the values each reference ``from_config`` reads)."""
    check_supported(cfg)
    M, rec = cfg.MODEL, cfg.MODEL.ROI_RECOGNIZER_HEAD
    ag = M.ANCHOR_GENERATOR
    kw: Dict[str, Any] = dict(
        pixel_mean=tuple(M.PIXEL_MEAN), pixel_std=tuple(M.PIXEL_STD),
        mask_inference=bool(M.ROI_MASK_HEAD.MASK_INFERENCE),
        # MODEL.ORIENTATION_ON is not an argument: it decides whether the checkpoint HAS box_predictor.orientation_pred
        # (rotated_fast_rcnn.py:547-549), and B200GlassROIHeads follows the checkpoint (configs/glass_finetune_textocr.yaml
        # is the one shipped config without it).
        # d2 RPN.from_config / RotatedAnchorGenerator.from_config
        rpn_kwargs=dict(anchor_sizes=tuple(tuple(s) for s in ag.SIZES), anchor_ratios=tuple(ag.ASPECT_RATIOS[0]),
                        anchor_angles=tuple(ag.ANGLES[0]), bbox_reg_weights=tuple(M.RPN.BBOX_REG_WEIGHTS),
                        pre_nms_topk=int(M.RPN.PRE_NMS_TOPK_TEST), post_nms_topk=int(M.RPN.POST_NMS_TOPK_TEST),
                        nms_thresh=float(M.RPN.NMS_THRESH)),
        # recognizers_hybrid_head.py:184-217 (_init_box_head), rotated_fast_rcnn.py:556-585 (output layers)
        box_pooler_resolution=int(M.ROI_BOX_HEAD.POOLER_RESOLUTION),
        box_pooler_sampling_ratio=int(M.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO),
        box_reg_weights=tuple(M.ROI_BOX_HEAD.BBOX_REG_WEIGHTS),
        score_thresh=float(M.ROI_HEADS.SCORE_THRESH_TEST), nms_thresh=float(M.ROI_HEADS.NMS_THRESH_TEST),
        detections_per_image=int(cfg.TEST.DETECTIONS_PER_IMAGE),
        # recognizers_hybrid_head.py:444-469 (_init_recognizer_head): pooler [H, W], POOLER_SAMPLING_RATIO; ASTER_V2's
        # from_config (recognizer_decoder.py:76-84): classes = character set + [GO] + [s], MAX_WORD_LENGTH + 1 steps
        recog_pool=(int(rec.POOLER_RESOLUTION_HEIGHT), int(rec.POOLER_RESOLUTION_WIDTH)),
        recog_sampling_ratio=int(rec.POOLER_SAMPLING_RATIO),
        num_text_classes=len(rec.CHARACTER_SET) + 2, max_word_len=int(rec.MAX_WORD_LENGTH) + 1,
    )
    if M.META_ARCHITECTURE == "GlassRCNN":          # glass_rcnn.py:37-55 (hasattr probes on cfg.POST_PROCESSING)
        pp = cfg.POST_PROCESSING
        kw.update(filter_small_boxes=pp.get("MIN_BOX_DIMENSION"), inflate_ratio=pp.get("INFLATE_RATIO"),
                  drop_overlapping_boxes=pp.get("DROP_OVERLAPPING"))
    return kw


def _state_dict(cfg: CfgNode, state_dict_or_path):
    if isinstance(state_dict_or_path, (str, os.PathLike)):
        from .weights import load_checkpoint
        return load_checkpoint(str(state_dict_or_path), mask=bool(cfg.MODEL.ROI_MASK_HEAD.MASK_INFERENCE))
    return state_dict_or_path


def build_model(cfg: CfgNode, state_dict, device="cuda"):
    """detectron2 ``build_model(cfg)`` + ``DetectionCheckpointer.load``, in one step (weights are packed at
    construction).  ``state_dict`` is a d2-style state_dict or the path of a ``.pth`` checkpoint."""
    from .modeling.glass_rcnn import B200GlassRCNN
    return B200GlassRCNN(_state_dict(cfg, state_dict), device=device, **model_kwargs(cfg))


def post_processing_config(cfg: CfgNode):
    from .postprocess import PostProcessingConfig
    pp = cfg.POST_PROCESSING
    return PostProcessingConfig(SKIP_ALL=pp.SKIP_ALL, MIN_BOX_DIMENSION=pp.MIN_BOX_DIMENSION,
                                MERGE_IOA_THRESH=pp.MERGE_IOA_THRESH,
                                PAIRS_HEIGHT_RATIO_THRESH=pp.PAIRS_HEIGHT_RATIO_THRESH,
                                VALID_CONFIDENCE=pp.VALID_CONFIDENCE, DETECT_THRESHOLD=pp.DETECT_THRESHOLD,
                                TEXT_THRESHOLD=pp.TEXT_THRESHOLD, MAX_ANGLE_DIFF=pp.MAX_ANGLE_DIFF)


def build_post_processor(cfg: CfgNode):
    """glass/postprocess/post_processor_rotated_boxes.py:23-29: the class named by cfg.POST_PROCESSING.NAME."""
    from .postprocess import B200PostProcessor
    name = cfg.POST_PROCESSING.NAME
    if name == "PostProcessorAcademic":
        return B200PostProcessor(post_processing_config(cfg), text_filter=True)
    if name == "PostProcessorRotatedBoxes":
        return B200PostProcessor(post_processing_config(cfg), text_filter=False)
    raise UnsupportedConfig(f"POST_PROCESSING.NAME = {name!r}: PostProcessorAcademic / PostProcessorRotatedBoxes only")


def build_runner(cfg: CfgNode, state_dict, device="cuda", post_process: bool = True):
    """GlassRunner(model_path, config_path, post_process) (glass/inference/glass_runner.py:24-70)."""
    from .runner import B200GlassRunner
    from .text import TextDecoder
    runner = B200GlassRunner(_state_dict(cfg, state_dict), min_target_size=int(cfg.INPUT.MIN_SIZE_TEST),
                             max_target_size=int(cfg.INPUT.MAX_SIZE_TEST),
                             max_upscale_ratio=float(cfg.INPUT.MAX_UPSCALE_RATIO), input_format=cfg.INPUT.FORMAT,
                             device=device, post_processor=build_post_processor(cfg) if post_process else None,
                             **model_kwargs(cfg))
    runner.text_decoder = TextDecoder(cfg.MODEL.ROI_RECOGNIZER_HEAD.CHARACTER_SET)
    return runner
