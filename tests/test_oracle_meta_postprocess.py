"""a17 -- GlassRCNN._postprocess (glass_rcnn.py:103-128) against golden vectors written by the reference's OWN
``_postprocess`` / ``filter_small_boxes`` / ``resize_boxes`` / ``detector_postprocess``
(tools/make_golden_meta_postprocess.py -> tests/golden/meta_postprocess.pt).  Two arms: the oracle's restatement
(oracle/postprocess.py) and the product's host glue (modeling/glass_rcnn.py; index glue on <= 100 boxes, runs on the
device the boxes live on -- here CPU tensors, no kernel involved)."""
import os

import pytest
import torch

from golden_common import make_meta_postprocess_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "meta_postprocess.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


@pytest.mark.parametrize("i", range(7))
def test_oracle_matches_reference(golden, i):
    from oracle import postprocess as pp
    c = golden["cases"][i]
    boxes, _ = make_meta_postprocess_inputs(c["seed"], c["n"], c["hw"])
    oh, ow = c["out_hw"] or c["hw"]
    out, idx = pp.glass_rcnn_postprocess(boxes, c["hw"], oh, ow, c["min_box_dim"], c["inflate_ratio"])
    assert torch.equal(idx, c["idx"])
    assert torch.equal(out, c["boxes"])


@pytest.mark.parametrize("i", range(7))
def test_product_postprocess_matches_reference(golden, i):
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    from glass_text_spotting_b200.structures import Instances, RotatedBoxes
    c = golden["cases"][i]
    boxes, scores = make_meta_postprocess_inputs(c["seed"], c["n"], c["hw"])
    inst = Instances(c["hw"], pred_boxes=RotatedBoxes(boxes.clone()), scores=scores, orig_idx=torch.arange(c["n"]))
    model = B200GlassRCNN.__new__(B200GlassRCNN)          # host glue only: no weights, no device
    model.filter_small_boxes, model.inflate_ratio = c["min_box_dim"], c["inflate_ratio"]
    inp = {} if c["out_hw"] is None else {"height": c["out_hw"][0], "width": c["out_hw"][1]}
    out = model._postprocess([inst], [inp], [c["hw"]])[0]["instances"]
    assert tuple(out.image_size) == c["image_size"]
    assert torch.equal(out.orig_idx, c["idx"])
    assert torch.equal(out.pred_boxes.tensor, c["boxes"])
    assert torch.equal(out.scores, c["scores"])


def test_cases_exercise_every_branch(golden):
    """Each filter must actually drop something, the clip must move something, inflation must widen something."""
    by = {c["seed"]: c for c in golden["cases"]}
    assert len(by[1]["idx"]) < len(by[0]["idx"]) < by[0]["n"]          # small-box filter on top of the empty filter
    b3, _ = make_meta_postprocess_inputs(3, by[3]["n"], by[3]["hw"])
    sx = by[3]["out_hw"][1] / by[3]["hw"][1]
    plain = b3[by[3]["idx"]][:, 2] * sx
    flat = b3[by[3]["idx"]][:, 4].abs() < 1e-3
    assert flat.any() or (by[3]["boxes"][:, 2] > 0).all()
    assert (by[3]["boxes"][:, 2][~flat] != plain[~flat]).any()


def test_detector_postprocess_alias_and_proposals(golden):
    from glass_text_spotting_b200.modeling.glass_rcnn import detector_postprocess
    from glass_text_spotting_b200.structures import Instances, RotatedBoxes
    from oracle import postprocess as pp
    a = golden["alias"]
    boxes, scores = make_meta_postprocess_inputs(a["seed"], a["n"], a["hw"])
    # oracle
    ob, keep, orb = pp.detector_postprocess(boxes, a["hw"], *a["out_hw"], rboxes_alias=True)
    assert torch.equal(torch.arange(a["n"])[keep], a["idx"]) and torch.equal(ob, a["boxes"]) and torch.equal(orb, a["rboxes"])
    assert not torch.equal(a["boxes"], a["rboxes"])      # the reference really scales the alias twice
    # product
    inst = Instances(a["hw"], pred_boxes=RotatedBoxes(boxes.clone()), scores=scores, orig_idx=torch.arange(a["n"]))
    inst.pred_rboxes = inst.pred_boxes
    out = detector_postprocess(inst, *a["out_hw"])
    assert torch.equal(out.orig_idx, a["idx"]) and torch.equal(out.pred_boxes.tensor, a["boxes"])
    assert torch.equal(out.pred_rboxes.tensor, a["rboxes"])
    assert torch.equal(inst.pred_boxes.tensor, boxes)    # the caller's boxes are left alone
    # an independent pred_rboxes is scaled once
    inst.pred_rboxes = RotatedBoxes(boxes.clone())
    out = detector_postprocess(inst, *a["out_hw"])
    assert torch.equal(out.pred_rboxes.tensor, a["boxes"])
    # proposal_boxes only (:151-154)
    p = golden["proposals"]
    inst = Instances(p["hw"], proposal_boxes=RotatedBoxes(boxes.clone()), objectness_logits=scores, orig_idx=torch.arange(p["n"]))
    out = detector_postprocess(inst, *p["out_hw"])
    assert torch.equal(out.orig_idx, p["idx"]) and torch.equal(out.proposal_boxes.tensor, p["boxes"])


def test_drop_overlapping_is_refused():
    """cfg.POST_PROCESSING.DROP_OVERLAPPING crashes in the reference (RotatedBoxes has no .shape); we refuse it up front."""
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    with pytest.raises(NotImplementedError):
        B200GlassRCNN({}, drop_overlapping_boxes=True)
