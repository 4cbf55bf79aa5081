"""The oracle's box-branch inference (oracle/model.py GlassOracle.box_inference) against golden vectors from the
reference's OWN RotatedFastRCNNOutputs.inference / fast_rcnn_inference_single_image_rotated
(tools/make_golden_box_inference.py -> tests/golden/box_inference.pt)."""
import os

import pytest
import torch

from golden_common import make_box_inference_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "box_inference.pt")


@pytest.mark.parametrize("i", range(4))
def test_box_inference_matches_reference(i):
    from oracle import model as om
    c = torch.load(GOLDEN, weights_only=False)["cases"][i]
    logits, deltas, orient, proposals = make_box_inference_inputs(c["seed"], c["r"], c["hw"])
    o = om.GlassOracle(om.HotPathConfig())
    with torch.no_grad():
        det = o.box_inference(logits, deltas, orient, proposals, c["hw"])
    assert torch.equal(det["kept_proposal_idx"], c["kept"])
    assert torch.equal(det["pred_boxes"], c["pred_boxes"])
    assert torch.equal(det["scores"], c["scores"])
    assert torch.equal(det["pred_classes"], c["pred_classes"])
    assert torch.equal(det["orientations"], c["orientations"])
