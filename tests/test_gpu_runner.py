"""GPU tests of the pre-processing row: device bilinear resize vs
torch.nn.functional.interpolate (what GlassRunner._image_to_tensor calls), and the runner end to end."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,out", [((97, 130), (200, 260)), ((300, 200), (150, 111)), ((64, 64), (64, 64)),
                                       ((1024, 1024), (1200, 1200))])
def test_resize_matches_torch_interpolate(glass_lib, shape, out):
    from glass_text_spotting_b200 import ops
    rng = np.random.RandomState(shape[0] + out[1])
    img = rng.randint(0, 256, size=shape + (3,), dtype=np.uint8)
    got = ops.resize_bilinear_u8(torch.as_tensor(img).cuda(), out, flip_channels=False).cpu()
    x = torch.as_tensor(img.transpose(2, 0, 1)).float().unsqueeze(0)
    ref = F.interpolate(x, size=out, mode="bilinear", align_corners=False)[0] if out != shape else x[0]
    assert torch.allclose(got, ref, rtol=1e-5, atol=2e-4), (got - ref).abs().max()
    flipped = ops.resize_bilinear_u8(torch.as_tensor(img).cuda(), out, flip_channels=True).cpu()
    assert torch.equal(flipped, got.flip(0))


def test_runner_end_to_end(glass_lib):
    from glass_text_spotting_b200 import weights
    from glass_text_spotting_b200.runner import B200GlassRunner
    runner = B200GlassRunner(weights.random_state_dict(0), min_target_size=320, max_target_size=384, detections_per_image=5)
    img = np.random.RandomState(0).randint(0, 256, size=(200, 260, 3), dtype=np.uint8)
    preds = runner(img)
    assert preds.image_size == (200, 260) and len(preds) <= 5
    assert tuple(preds.pred_text_prob.shape[1:]) == (26, 97)
    words = runner.read_text(preds)
    assert len(words) == len(preds) and all(isinstance(w["text"], str) for w in words)
    if len(preds):
        assert float(preds.pred_boxes.tensor[:, 0].max()) < 260 * 1.5   # boxes are in original-image coordinates
