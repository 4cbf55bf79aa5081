"""GPU parity of the mask branch: mask head on glass_conv_gemm vs the torch layers of oracle/mask.py,
the rotated paste vs the oracle AND vs golden vectors written by the reference's own paste_masks_in_image
(tests/golden/paste_masks.pt), and the branch wired into B200GlassRCNN.inference."""
import os

import numpy as np
import pytest
import torch

from golden_common import make_paste_inputs
from parity_common import close

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "paste_masks.pt")


def _rois(g, n, img=256.0, batch=2):
    import math
    cx, cy = torch.rand(n, generator=g) * img, torch.rand(n, generator=g) * img
    w = torch.exp(torch.rand(n, generator=g) * (math.log(200) - math.log(12)) + math.log(12))
    h = w * (0.15 + 0.85 * torch.rand(n, generator=g))
    a = torch.rand(n, generator=g) * 360 - 180
    b = torch.sort(torch.randint(0, batch, (n,), generator=g).float())[0]
    return torch.stack((b, cx, cy, w, h, a), 1).contiguous()


def test_mask_head_matches_oracle(glass_lib):
    from glass_text_spotting_b200 import ops
    from glass_text_spotting_b200.modeling.mask_head import B200MaskHead
    from oracle import d2_ops, mask as om
    g = torch.Generator().manual_seed(41)
    sizes, scales = [64, 32, 16, 8, 4], [1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64]
    acts = {f"p{i + 2}": ops.Act.from_nchw(torch.randn(2, 256, s, s, generator=g).cuda()) for i, s in enumerate(sizes)}
    feats = [acts[f"p{i + 2}"].to_nchw().cpu() for i in range(5)]
    rois = _rois(g, 23)
    head = om.seeded_mask_head(3)
    with torch.no_grad():
        pooled = d2_ops.roi_pooler(feats, [rois[rois[:, 0] == b][:, 1:] for b in range(2)], (14, 14), scales, 0)
        want = head(pooled)
    sd = {"roi_heads.mask_head." + k: v for k, v in head.state_dict().items()}
    got = B200MaskHead(sd)(acts, rois.cuda())
    assert tuple(got.shape) == (23, 1, 28, 28)
    close(got, want, "pred_masks", rtol=1e-3, atol=1e-4)
    # no detections
    assert tuple(B200MaskHead(sd)(acts, torch.zeros((0, 6)).cuda()).shape) == (0, 1, 28, 28)


@pytest.mark.parametrize("i", range(4))
def test_paste_matches_reference_golden(glass_lib, i):
    from glass_text_spotting_b200.modeling.mask_head import B200MaskHead
    from oracle import mask as om
    c = torch.load(GOLDEN, weights_only=False)["cases"][i]
    h, w = c["hw"]
    masks, boxes = make_paste_inputs(c["seed"], c["n"], h, w)
    out, soft = B200MaskHead.paste(masks.cuda(), boxes.cuda(), (h, w), 0.5, want_soft=True)
    assert tuple(out.shape) == (c["n"], h, w) and out.dtype == torch.bool
    if c["n"] == 0:
        return
    ref_soft = om.do_paste_mask_rotated(masks[:, None], boxes, h, w)
    assert (soft.cpu() - ref_soft).abs().max().item() < 2e-5
    assert torch.allclose(soft.cpu()[:, ::7, ::5], c["soft_sample"], atol=2e-5)
    ref = np.unpackbits(c["packed"].numpy())[: c["n"] * h * w].reshape(c["n"], h, w).astype(bool)
    diff = out.cpu().numpy() != ref
    # a pixel can only flip where the sampled value sits within rounding of the threshold
    assert diff.sum() <= 2 and ((ref_soft.numpy()[diff] - 0.5).__abs__() < 2e-5).all(), int(diff.sum())
    assert abs(int(out.sum()) - c["count"]) <= 2


def test_paste_full_size_properties(glass_lib):
    """1024 x 1024, 100 boxes: an all-ones mask pastes to exactly the pixels whose centres lie inside the rotated box
    (up to the half-texel bilinear falloff at the rim); a zero mask pastes to nothing."""
    from glass_text_spotting_b200.modeling.mask_head import B200MaskHead
    g = torch.Generator().manual_seed(5)
    n = 100
    boxes = torch.stack([torch.rand(n, generator=g) * 1024, torch.rand(n, generator=g) * 1024,
                         torch.rand(n, generator=g) * 300 + 20, torch.rand(n, generator=g) * 80 + 10,
                         torch.rand(n, generator=g) * 360 - 180], 1)
    ones = torch.ones(n, 28, 28)
    out = B200MaskHead.paste(ones.cuda(), boxes.cuda(), (1024, 1024), 0.5)
    assert not B200MaskHead.paste(torch.zeros(n, 28, 28).cuda(), boxes.cuda(), (1024, 1024), 0.5).any()
    ys, xs = torch.meshgrid(torch.arange(1024.0) + 0.5, torch.arange(1024.0) + 0.5, indexing="ij")
    for i in (0, 17, 63):
        cx, cy, w, h, a = boxes[i].tolist()
        t = torch.deg2rad(torch.tensor(a))
        gx, gy = xs - cx, ys - cy
        rx = gx * torch.cos(t) - gy * torch.sin(t)
        ry = gx * torch.sin(t) + gy * torch.cos(t)
        # the sampled value is ramp(u) * ramp(v): 1 up to half a texel from the box edge, 0.5 exactly at the edge (zero
        # padding outside), so the core of the box is set, everything beyond the edge is clear, corners are cut
        inside = (rx.abs() < w / 2 * (1 - 1.02 / 28)) & (ry.abs() < h / 2 * (1 - 1.02 / 28))
        outside = (rx.abs() > w / 2 * (1 + 1 / 28 * 0.02)) | (ry.abs() > h / 2 * (1 + 1 / 28 * 0.02))
        got = out[i].cpu()
        assert got[inside].all() and not got[outside].any()


def test_mask_inference_end_to_end(glass_lib):
    """MASK_INFERENCE on: Instances carry pred_masks pasted at the output size, consistent with their boxes."""
    from glass_text_spotting_b200 import weights
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    sd = weights.random_state_dict(0)
    sd.update(weights.random_mask_head_state_dict(0))
    model = B200GlassRCNN(sd, detections_per_image=6, mask_inference=True)
    img = torch.randint(0, 256, (3, 160, 224), generator=torch.Generator().manual_seed(3)).float()
    out = model([{"image": img, "height": 320, "width": 448}, {"image": img}])
    for r, (hh, ww) in zip(out, [(320, 448), (160, 224)]):
        inst = r["instances"]
        assert inst.pred_masks.dtype == torch.bool and tuple(inst.pred_masks.shape) == (len(inst), hh, ww)
    # without the flag nothing changes
    plain = B200GlassRCNN(sd, detections_per_image=6)([{"image": img}])[0]["instances"]
    assert not plain.has("pred_masks")
    assert torch.equal(plain.pred_boxes.tensor, out[1]["instances"].pred_boxes.tensor)
