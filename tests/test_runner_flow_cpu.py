"""B200GlassRunner's host flow against golden vectors written by the reference's OWN ``GlassRunner`` methods
(tools/make_golden_runner.py -> tests/golden/runner.pt): the scale-ratio rule, the tensor the model is handed (size,
channel handling, greyscale), the scale-back of the boxes, ``_image_size`` and the post-processor hand-off
(glass/inference/glass_runner.py:72-148).  The device resize kernel is stood in for by ``F.interpolate`` -- the call
the reference makes, and what tests/test_gpu_runner.py equates the kernel with on the GPU."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from golden_common import make_runner_case

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "runner.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def _runner(fmt="BGR"):
    from glass_text_spotting_b200.runner import B200GlassRunner
    r = B200GlassRunner.__new__(B200GlassRunner)      # host flow only: no weights, no device
    r.min_target_size, r.max_target_size, r.max_upscale_ratio, r.input_format, r.device = 1200, 1600, 2, fmt, "cpu"
    return r


def test_scale_ratio_rule(golden):
    r = _runner()
    assert [float(r.get_inference_scale_ratio(s)) for s in golden["shapes"]] == golden["ratios"]


@pytest.mark.parametrize("i", [0, 2, 3])
def test_flow_matches_reference(golden, i, monkeypatch):
    from glass_text_spotting_b200 import ops
    from glass_text_spotting_b200.structures import Instances, RotatedBoxes
    c = golden["cases"][i]
    image, boxes, scores = make_runner_case(c["seed"], c["hw"])

    def resize_cpu(src_hwc, out_hw, flip_channels=False):
        x = src_hwc.permute(2, 0, 1).float()
        if flip_channels:
            x = x.flip(0)
        if tuple(out_hw) == tuple(x.shape[1:]):
            return x.clone()
        return F.interpolate(x[None], size=tuple(out_hw), mode="bilinear", align_corners=False)[0]
    monkeypatch.setattr(ops, "resize_bilinear_u8", resize_cpu)
    seen = {}

    def model(inputs):
        seen.update(inputs[0])
        sc = inputs[0]["image"].shape[1] / c["hw"][0]
        return [{"instances": Instances((inputs[0]["height"], inputs[0]["width"]),
                                        pred_boxes=RotatedBoxes(boxes.clone() * torch.tensor([sc, sc, sc, sc, 1.0])),
                                        scores=scores.clone())}]

    def post(preds):
        seen["post_in_size"], seen["post_in_boxes"] = tuple(preds.image_size), preds.pred_boxes.tensor.clone()
        return preds[preds.scores > 0.5]
    r = _runner(c["format"])
    r.model, r.post_processor = model, post
    out = r(image)
    t = seen["image"]
    assert tuple(t.shape) == c["tensor_shape"] and (seen["height"], seen["width"]) == c["model_hw"]
    assert torch.allclose(t[:, ::7, ::5], c["tensor_sample"], rtol=0, atol=1e-4)
    assert abs(float(t.double().sum()) - c["tensor_sum"]) <= 1e-6 * abs(c["tensor_sum"])
    assert seen["post_in_size"] == c["post_in_size"] and torch.equal(seen["post_in_boxes"], c["post_in_boxes"])
    assert tuple(out.image_size) == c["out_size"]
    assert torch.equal(out.pred_boxes.tensor, c["out_boxes"]) and torch.equal(out.scores, c["out_scores"])


def test_rgb_is_a_channel_flip_where_the_reference_crashes(golden, monkeypatch):
    """input_format "RGB": the reference hands torch.as_tensor a negatively strided view and raises (golden case 1);
    here the flip is fused into the resize kernel's channel index."""
    from glass_text_spotting_b200 import ops
    assert "raises" in golden["cases"][1]
    image, _, _ = make_runner_case(1, (150, 111))
    calls = {}
    monkeypatch.setattr(ops, "resize_bilinear_u8", lambda src, out_hw, flip_channels=False: calls.update(flip=flip_channels, hw=tuple(out_hw)) or src)
    r = _runner("RGB")
    _, scale = r.image_to_tensor(image)
    assert calls == {"flip": True, "hw": (300, 222)} and scale == 2
    with pytest.raises(AssertionError):
        from glass_text_spotting_b200.runner import B200GlassRunner
        B200GlassRunner({}, input_format="YUV")


def test_grey_conversion_is_the_reference_expression():
    rng = np.random.RandomState(5)
    img = rng.randint(0, 256, size=(9, 11, 3), dtype=np.uint8)
    want = np.uint8(0.2125 * img[:, :, 0] + 0.7154 * img[:, :, 1] + 0.0721 * img[:, :, 2])
    r = _runner("GREY")
    import glass_text_spotting_b200.ops as ops
    orig = ops.resize_bilinear_u8
    try:
        ops.resize_bilinear_u8 = lambda src, out_hw, flip_channels=False: src
        t, _ = r.image_to_tensor(img)
    finally:
        ops.resize_bilinear_u8 = orig
    assert np.array_equal(t.numpy()[:, :, 0], want) and np.array_equal(t.numpy()[:, :, 2], want)
