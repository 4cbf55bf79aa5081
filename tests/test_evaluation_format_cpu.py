"""Evaluator wire format against golden vectors written by the reference's own
text_evaluator.py functions (tools/make_golden_eval_format.py -> tests/golden/eval_format.pt).  Host logic only."""
import os

import numpy as np
import pytest
import torch

from golden_common import make_eval_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "eval_format.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def _instances(c):
    from glass_text_spotting_b200.structures import Instances, RotatedBoxes
    boxes, scores, probs = make_eval_inputs(c["seed"], c["n"])
    assert torch.equal(boxes, c["boxes"]) and torch.equal(scores, c["scores"])
    return Instances((1024, 1024), pred_boxes=RotatedBoxes(boxes), scores=scores, pred_text_prob=probs), probs


@pytest.mark.parametrize("i", range(4))
def test_records_match_reference(golden, i):
    from glass_text_spotting_b200 import evaluation as ev
    from glass_text_spotting_b200.text import TextDecoder
    c = golden["cases"][i]
    inst, probs = _instances(c)
    dec = TextDecoder()
    for flag, key in ((True, "records"), (False, "records_keep_specials")):
        got = ev.instances_to_coco_json(inst, c["image_id"], dec, flag)
        ref = c[key]
        assert len(got) == len(ref)
        for g, r in zip(got, ref):
            assert set(g) == set(r)
            for k in ("image_id", "category_id", "rec", "rboxes"):
                assert g[k] == r[k], k
            for k in ("polys", "boxes"):
                assert np.allclose(np.array(g[k]), np.array(r[k]), rtol=0, atol=1e-9), k
            assert abs(g["score_text"] - r["score_text"]) <= 1e-6 * r["score_text"] + 1e-12
            assert g["score_detection"] == r["score_detection"]
            assert g["character_probs"] == np.float64(probs[r["character_probs"]["det_index"]].numpy()).tolist()
    if c["n"]:
        texts, scores, _ = ev.get_instances_text(probs, dec, True)
        assert texts == c["texts"]
        assert np.allclose(scores, c["text_scores"], rtol=1e-6, atol=1e-12)


def test_polygons_match_reference(golden):
    from glass_text_spotting_b200 import evaluation as ev
    boxes = golden["cases"][1]["boxes"].numpy()
    assert np.array_equal(ev.rotated_boxes_to_polygons(boxes), golden["polygons_case1"].numpy())
    assert ev.rotated_boxes_to_polygons(np.zeros((0, 5))).shape == (0, 4, 2)
    assert ev.boxes_to_polygons(np.zeros((0, 4))).shape == (0, 4, 2)


@pytest.mark.parametrize("dataset", ["totaltext", "icdar15"])
def test_eval_files_match_reference(golden, dataset):
    from glass_text_spotting_b200 import evaluation as ev
    records = [r for c in golden["cases"] for r in c["records"]]
    got = ev.to_eval_lines(records, dataset, golden["text_cf_th"], golden["detection_cf_th"])
    assert got == golden["eval_files"][dataset]
    with pytest.raises(ValueError):
        ev.eval_file_name(3, "coco")


def test_packed_record_round_trip():
    """pack_detections' layout -> Instances -> records (what rank 0 does with the all-gathered tensor)."""
    from glass_text_spotting_b200 import evaluation as ev
    from glass_text_spotting_b200.text import TextDecoder
    boxes, scores, probs = make_eval_inputs(7, 5)
    rec = torch.zeros(2, 8, 10 + 26 * 97)
    rec[1, :5, 0] = 1.0
    rec[1, :5, 1:6], rec[1, :5, 6], rec[1, :5, 10:] = boxes, scores, probs.reshape(5, -1)
    insts = ev.instances_from_packed(rec, [(1024, 1024), (512, 640)])
    assert len(insts[0]) == 0 and len(insts[1]) == 5 and insts[1].image_size == (512, 640)
    assert torch.equal(insts[1].pred_boxes.tensor, boxes) and torch.equal(insts[1].pred_text_prob, probs)
    assert ev.instances_to_coco_json(insts[0], 1, TextDecoder()) == []
    assert len(ev.instances_to_coco_json(insts[1], 2, TextDecoder())) <= 5
