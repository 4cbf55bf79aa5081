"""oracle/mask.py's rotated paste against golden vectors from the reference's own paste_masks_in_image
(tools/make_golden_paste.py -> tests/golden/paste_masks.pt)."""
import os

import numpy as np
import pytest
import torch

from golden_common import make_paste_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "paste_masks.pt")


@pytest.mark.parametrize("i", range(4))
def test_paste_matches_reference(i):
    from oracle import mask as om
    c = torch.load(GOLDEN, weights_only=False)["cases"][i]
    h, w = c["hw"]
    masks, boxes = make_paste_inputs(c["seed"], c["n"], h, w)
    out = om.paste_masks_in_image(masks, boxes, (h, w), 0.5)
    assert tuple(out.shape) == (c["n"], h, w)
    if c["n"] == 0:
        assert out.dtype == torch.uint8
        return
    ref = np.unpackbits(c["packed"].numpy())[: c["n"] * h * w].reshape(c["n"], h, w).astype(bool)
    assert np.array_equal(out.numpy(), ref)
    assert int(out.sum()) == c["count"]
    soft = om.do_paste_mask_rotated(masks[:, None], boxes, h, w)
    assert torch.equal(soft[:, ::7, ::5], c["soft_sample"])
    assert abs(float(soft.double().sum()) - c["soft_sum"]) < 1e-6 * max(1.0, abs(c["soft_sum"]))


def test_mask_head_shapes_and_range():
    from oracle import mask as om
    m = om.seeded_mask_head(0)
    x = torch.randn(3, 256, 14, 14)
    with torch.no_grad():
        y = m(x)
    assert tuple(y.shape) == (3, 1, 28, 28) and float(y.min()) >= 0 and float(y.max()) <= 1
    assert float(y.std()) > 0.05   # the seeded head is not degenerate
    keys = set(m.state_dict())
    assert {"mask_fcn1.weight", "mask_fcn4.bias", "deconv.weight", "predictor.bias"} <= keys
