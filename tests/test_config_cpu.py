"""glass_text_spotting_b200.config: the reference's YAML configs -> constructor arguments (host logic, no GPU).
The shipped configs are read from /root/reference when it exists (authoring container; the expected values below were
checked against them); an inline config with the same keys as configs/glass_pretrain.yaml always runs."""
import os

import pytest

REF_CONFIGS = "/root/reference/configs"
PRETRAIN_LIKE = {   # the inference-relevant keys of configs/glass_pretrain.yaml:1-100, 136-141
    "MODEL": {
        "ROTATED_BOXES_ON": True, "ORIENTATION_ON": True, "RECOGNIZER_ON": True, "MASK_ON": True,
        "ROI_RECOGNIZER_HEAD": {"NAME": "RecognizerRCNNHeadV3", "MAX_WORD_LENGTH": 25, "POOLER_RESOLUTION_WIDTH": 32,
                                "POOLER_RESOLUTION_HEIGHT": 8, "POOLER_TYPE": "ROIAlignRotated", "NORM": "SyncBN",
                                "IN_FEATURES": ["p2", "p3"], "SAMPLING_RATIO": 0,
                                "RECOGNIZER_HEAD": {"BACKBONE": {"NAME": "CNN_V1_1"},
                                                    "ENCODER": {"NAME": "BiLSTMBlockV2", "NUM_OF_LAYERS": 2},
                                                    "DECODER": {"NAME": "ASTER_V2"}}},
        "ROI_MASK_HEAD": {"NAME": "RotatedMaskRCNNConvUpsampleHead", "NUM_CONV": 4, "POOLER_RESOLUTION": 14,
                          "POOLER_TYPE": "ROIAlignRotated"},
        "META_ARCHITECTURE": "GeneralizedRCNN",
        "BACKBONE": {"NAME": "build_resnet_fpn_backbone", "FREEZE_AT": 0},
        "RESNETS": {"OUT_FEATURES": ["res2", "res3", "res4", "res5"], "DEPTH": 50, "NORM": "SyncBN"},
        "FPN": {"IN_FEATURES": ["res2", "res3", "res4", "res5"], "OUT_CHANNELS": 256, "NORM": "SyncBN"},
        "ANCHOR_GENERATOR": {"NAME": "RotatedAnchorGenerator", "SIZES": [[16], [32], [64], [128], [256]],
                             "ASPECT_RATIOS": [[0.2, 0.5, 1.0]], "ANGLES": [[-90, -45, 0, 45]]},
        "PROPOSAL_GENERATOR": {"NAME": "RotatedRPN"},
        "RPN": {"IN_FEATURES": ["p2", "p3", "p4", "p5", "p6"], "PRE_NMS_TOPK_TEST": 1000,
                "BBOX_REG_WEIGHTS": [1.0, 1.0, 1.0, 1.0, 2.0], "POST_NMS_TOPK_TEST": 100},
        "ROI_HEADS": {"NAME": "MaskRotatedRecognizerHybridHead", "IN_FEATURES": ["p2", "p3", "p4", "p5", "p6"],
                      "NUM_CLASSES": 1, "NMS_THRESH_TEST": 0.35},
        "LOCAL_FEATURE_EXTRACTOR": {"NAME": "ResNetFeatureExtractor", "NUM_FEATURES": 256},
        "HYBRID_FUSION": {"NAME": "MultiAspectGCAttention", "NUM_FEATURES": 256},
        "ROI_BOX_HEAD": {"NAME": "FastRCNNConvFCHead", "NUM_FC": 2, "FC_DIM": 2048, "POOLER_RESOLUTION": 7,
                         "POOLER_SAMPLING_RATIO": 2, "POOLER_TYPE": "ROIAlignRotated", "NUM_CONV": 0,
                         "BBOX_REG_WEIGHTS": [10.0, 10.0, 5.0, 5.0, 10.0], "NORM": "SyncBN"},
    },
    "INPUT": {"MIN_SIZE_TEST": 1200, "MAX_SIZE_TEST": 1600},
}

EXPECTED = dict(
    pixel_mean=(103.530, 116.280, 123.675), pixel_std=(1.0, 1.0, 1.0), mask_inference=False,
    rpn_kwargs=dict(anchor_sizes=((16,), (32,), (64,), (128,), (256,)), anchor_ratios=(0.2, 0.5, 1.0),
                    anchor_angles=(-90, -45, 0, 45), bbox_reg_weights=(1.0, 1.0, 1.0, 1.0, 2.0), pre_nms_topk=1000,
                    post_nms_topk=100, nms_thresh=0.7),
    box_pooler_resolution=7, box_pooler_sampling_ratio=2, box_reg_weights=(10.0, 10.0, 5.0, 5.0, 10.0),
    score_thresh=0.05, nms_thresh=0.35, detections_per_image=100, recog_pool=(8, 32), recog_sampling_ratio=0,
    num_text_classes=97, max_word_len=26)


def test_inline_pretrain_config_maps_to_the_constructor_defaults():
    """The mapped values ARE the defaults the modules were built and parity-tested with."""
    import inspect
    from glass_text_spotting_b200 import config
    from glass_text_spotting_b200.modeling.roi_heads import B200GlassROIHeads
    from glass_text_spotting_b200.modeling.rpn import B200RotatedRPN
    cfg = config.load_config(PRETRAIN_LIKE)
    kw = config.model_kwargs(cfg)
    assert kw == EXPECTED
    heads = inspect.signature(B200GlassROIHeads.__init__).parameters
    for k in ("box_pooler_resolution", "box_pooler_sampling_ratio", "box_reg_weights", "score_thresh", "nms_thresh",
              "detections_per_image", "recog_pool", "recog_sampling_ratio", "num_text_classes", "max_word_len"):
        assert heads[k].default == kw[k], k
    rpn = inspect.signature(B200RotatedRPN.__init__).parameters
    for k, v in kw["rpn_kwargs"].items():
        assert tuple(rpn[k].default) == tuple(v) if isinstance(v, tuple) else rpn[k].default == v, k
    assert cfg.INPUT.MAX_UPSCALE_RATIO == 2 and cfg.INPUT.FORMAT == "BGR"
    assert not hasattr(cfg.POST_PROCESSING, "INFLATE_RATIO") and hasattr(cfg.POST_PROCESSING, "MIN_BOX_DIMENSION")


@pytest.mark.skipif(not os.path.isdir(REF_CONFIGS), reason="reference configs only exist in the authoring container")
@pytest.mark.parametrize("name,meta", [("glass_pretrain.yaml", "GeneralizedRCNN"), ("glass_finetune_totaltext.yaml", "GlassRCNN"),
                                       ("glass_finetune_icdar15.yaml", "GlassRCNN"), ("glass_finetune_textocr.yaml", "GlassRCNN")])
def test_shipped_configs(name, meta):
    from glass_text_spotting_b200 import config
    cfg = config.load_config(os.path.join(REF_CONFIGS, name))
    assert cfg.MODEL.META_ARCHITECTURE == meta
    kw = config.model_kwargs(cfg)
    want = dict(EXPECTED)
    if meta == "GlassRCNN":   # glass_rcnn.py:43-50: MIN_BOX_DIMENSION 2, no INFLATE_RATIO / DROP_OVERLAPPING key
        want.update(filter_small_boxes=2, inflate_ratio=None, drop_overlapping_boxes=None)
    assert kw == want
    pp = config.post_processing_config(cfg)
    assert (pp.MIN_BOX_DIMENSION, pp.VALID_CONFIDENCE, pp.DETECT_THRESHOLD, pp.MERGE_IOA_THRESH) == (2, 0.15, 0.25, 0.3)
    assert (cfg.INPUT.MIN_SIZE_TEST, cfg.INPUT.MAX_SIZE_TEST) == (1200, 1600)


@pytest.mark.parametrize("path,value", [("MODEL.RESNETS.DEPTH", 101), ("MODEL.BACKBONE.NAME", "build_resnet_backbone"),
                                        ("MODEL.PROPOSAL_GENERATOR.NAME", "RPN"), ("MODEL.FPN.NORM", ""),
                                        ("MODEL.ROI_RECOGNIZER_HEAD.RECOGNIZER_HEAD.DECODER.NAME", "ASTER"),
                                        ("MODEL.ROI_BOX_HEAD.NUM_FC", 1), ("MODEL.HYBRID_FUSION.HEADERS", 4)])
def test_other_architectures_are_refused_with_the_key(path, value):
    from glass_text_spotting_b200 import config
    cfg = config.load_config(PRETRAIN_LIKE)
    node = cfg
    parts = path.split(".")
    for p in parts[:-1]:
        node = node[p]
    node[parts[-1]] = value
    with pytest.raises(config.UnsupportedConfig, match=path.replace(".", r"\.")):
        config.model_kwargs(cfg)


def test_glass_rcnn_options_and_post_processor_selection():
    from glass_text_spotting_b200 import config
    from glass_text_spotting_b200.postprocess import B200PostProcessor
    cfg = config.load_config(PRETRAIN_LIKE)
    cfg.MODEL.META_ARCHITECTURE = "GlassRCNN"
    cfg.POST_PROCESSING.INFLATE_RATIO = 0.05
    cfg.POST_PROCESSING.MIN_BOX_DIMENSION = 3
    kw = config.model_kwargs(cfg)
    assert (kw["filter_small_boxes"], kw["inflate_ratio"], kw["drop_overlapping_boxes"]) == (3, 0.05, None)
    post = config.build_post_processor(cfg)
    assert isinstance(post, B200PostProcessor) and post.text_filter and post.cfg.MIN_BOX_DIMENSION == 3
    cfg.POST_PROCESSING.NAME = "PostProcessorRotatedBoxes"
    assert not config.build_post_processor(cfg).text_filter
    cfg.POST_PROCESSING.NAME = "Nope"
    with pytest.raises(config.UnsupportedConfig):
        config.build_post_processor(cfg)
    cfg.MODEL.ROI_MASK_HEAD.MASK_INFERENCE = True
    assert config.model_kwargs(cfg)["mask_inference"] is True
    with pytest.raises(NotImplementedError):
        config.load_config({"_BASE_": "x.yaml"})


@pytest.mark.parametrize("path,value,needle", [
    (("MODEL", "RPN", "PRE_NMS_TOPK_TEST"), 6000, "PRE_NMS_TOPK_TEST"),          # detectron2's default when the YAML omits it
    (("MODEL", "RPN", "POST_NMS_TOPK_TEST"), 1000, "POST_NMS_TOPK_TEST"),
    (("TEST", "DETECTIONS_PER_IMAGE"), 300, "DETECTIONS_PER_IMAGE"),
    (("MODEL", "ROI_RECOGNIZER_HEAD", "MAX_WORD_LENGTH"), 80, "MAX_WORD_LENGTH"),
    (("MODEL", "ROI_RECOGNIZER_HEAD", "POOLER_RESOLUTION_WIDTH"), 64, "POOLER_RESOLUTION_WIDTH"),
])
def test_kernel_limits_are_refused_by_key(path, value, needle):
    """the kernels' hard numeric limits are validated with the offending key named, not discovered at the
    first image as a C-side check or a bare assert."""
    import copy
    from glass_text_spotting_b200 import config
    raw = copy.deepcopy(PRETRAIN_LIKE)
    node = raw
    for k in path[:-1]:
        node = node.setdefault(k, {})
    node[path[-1]] = value
    with pytest.raises(config.UnsupportedConfig, match=needle):
        config.model_kwargs(config.load_config(raw))


def test_omitted_rpn_topk_inherits_d2_defaults_and_is_refused():
    import copy
    from glass_text_spotting_b200 import config
    raw = copy.deepcopy(PRETRAIN_LIKE)
    del raw["MODEL"]["RPN"]["PRE_NMS_TOPK_TEST"]
    with pytest.raises(config.UnsupportedConfig, match="PRE_NMS_TOPK_TEST = 6000"):
        config.check_supported(config.load_config(raw))
