"""Rows a5-a8 wired together: the oracle's ``box_branch`` (oracle/model.py) against golden vectors written by the
reference's OWN ``MaskRotatedRecognizerHybridHead._forward_box`` driving its own ``RotatedFastRCNNOutputLayers`` (built
by its ``from_config`` from the reference's pretrain config) and ``RotatedFastRCNNOutputs.inference``
(tools/make_golden_box_branch.py -> tests/golden/box_branch.pt)."""
import os

import pytest
import torch

from golden_common import make_box_branch_inputs, seeded_fill

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "box_branch.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


@pytest.mark.parametrize("i", range(3))
def test_box_branch_matches_reference(golden, i):
    from oracle import model as om
    c = golden["cases"][i]
    o = om.GlassOracle(om.HotPathConfig(max_detections_override=c["detections"]))
    rh = o.roi_heads
    assert sorted(rh.box_head.state_dict()) == c["head_keys"]
    assert sorted(rh.box_predictor.state_dict()) == c["predictor_keys"]
    seeded_fill(rh.box_head, 700 + c["seed"])
    seeded_fill(rh.box_predictor, 710 + c["seed"])
    with torch.no_grad():
        rh.box_predictor.cls_score.weight.mul_(4.0)
        rh.box_predictor.bbox_pred.weight.mul_(0.3)
        feats, proposals, hw = make_box_branch_inputs(c["seed"], c["r"])
        det = o.box_branch(feats, proposals, hw)
    # the values the reference's from_config hands its predictor are the oracle's configuration
    cfg = o.cfg
    assert c["thresholds"] == (cfg.score_thresh_test, cfg.nms_thresh_test, c["detections"], tuple(cfg.box_reg_weights))
    assert c["image_size"] == tuple(hw)
    k = len(c["scores"])
    assert len(det["scores"]) == k and (k == c["detections"] or k < c["r"])
    assert torch.equal(det["pred_classes"], c["pred_classes"])
    assert torch.allclose(det["pred_boxes"], c["pred_boxes"], rtol=1e-5, atol=1e-4)
    assert torch.allclose(det["scores"], c["scores"], rtol=1e-5, atol=1e-6)
    assert torch.equal(det["orientations"][:, 0], c["orientations"][:, 0])
    assert torch.allclose(det["orientations"][:, 1], c["orientations"][:, 1], rtol=1e-5, atol=1e-6)
    assert bool((c["scores"][:-1] >= c["scores"][1:]).all())
