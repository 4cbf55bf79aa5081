"""config.build_model / build_runner on the device: a model built from a reference-style config is the model built with
the constructors' defaults (the shipped configs ARE those defaults, tests/test_config_cpu.py), and the runner built
from the config runs the GlassRunner flow (glass/inference/glass_runner.py:24-109)."""
import numpy as np
import pytest
import torch

from test_config_cpu import PRETRAIN_LIKE

pytestmark = pytest.mark.gpu


def test_build_model_and_runner_from_config(glass_lib):
    from glass_text_spotting_b200 import config
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    from glass_text_spotting_b200.postprocess import B200PostProcessor
    from oracle import model as om
    K = 6
    img = om.synthetic_image(41, 160, 224)
    o = om.build_oracle(seed=7, calib_images=[img], cfg=om.HotPathConfig(max_detections_override=K))
    sd = o.state_dict()
    cfg = config.load_config(PRETRAIN_LIKE)
    cfg.TEST.DETECTIONS_PER_IMAGE = K
    built = config.build_model(cfg, sd)
    plain = B200GlassRCNN(sd, detections_per_image=K)
    assert built.filter_small_boxes is None and built.proposal_generator.post_nms_topk == 100
    a = built([{"image": img}])[0]["instances"]
    b = plain([{"image": img}])[0]["instances"]
    assert len(a) == len(b) > 0
    assert torch.equal(a.pred_boxes.tensor, b.pred_boxes.tensor) and torch.equal(a.pred_text_prob, b.pred_text_prob)
    # the GlassRCNN meta-architecture of the fine-tune configs adds the small-box filter (glass_rcnn.py:43-50, 117-118)
    cfg.MODEL.META_ARCHITECTURE = "GlassRCNN"
    assert config.model_kwargs(cfg)["filter_small_boxes"] == 2
    # GlassRunner from the same config: HWC uint8 image in, post-processed Instances out
    runner = config.build_runner(cfg, sd)
    assert isinstance(runner.post_processor, B200PostProcessor) and runner.model.filter_small_boxes == 2
    assert (runner.min_target_size, runner.max_target_size, runner.max_upscale_ratio) == (1200, 1600, 2.0)
    hwc = img.permute(1, 2, 0).to(torch.uint8).numpy()
    assert runner.get_inference_scale_ratio(hwc.shape) == 2.0          # 224 -> min(2, 1200/224)
    preds = runner(np.ascontiguousarray(hwc))
    assert preds.image_size == (160, 224) and preds.has("pred_boxes") and preds.has("pred_text_prob")
    assert preds.has("pred_polygons") and len(preds) <= K
    raw = config.build_runner(cfg, sd, post_process=False)(np.ascontiguousarray(hwc))
    assert not raw.has("pred_polygons") and len(raw) >= len(preds)
    texts = runner.read_text(preds)
    assert len(texts) == len(preds) and all(set(t) >= {"text", "score"} for t in texts)
