"""The oracle's restatement of the reference post-processor (oracle/postprocess.py) against golden vectors produced
by the reference's OWN classes (tools/make_golden_postprocess.py -> tests/golden/postprocess.pt)."""
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "postprocess.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def test_helpers_match_reference(golden):
    from oracle import postprocess as pp
    poly = pp.boxes_to_polygons(golden["helper_boxes"])
    assert torch.equal(poly, golden["helper_polygons"])
    back = pp.polygons_to_rotated_boxes(poly, golden["helper_boxes"][:, 4])
    assert torch.allclose(back, golden["helper_roundtrip"], atol=1e-4)


@pytest.mark.parametrize("i", range(8))
def test_post_process_matches_reference(golden, i):
    from oracle import postprocess as pp
    c = golden["cases"][i]
    # PostProcessorRotatedBoxes.__call__ (no text filter)
    boxes, idx, poly, iters = pp.post_process(c["boxes"], c["scores"], None)
    assert torch.equal(idx, c["rb_idx"]), (idx, c["rb_idx"])
    assert torch.allclose(boxes, c["rb_boxes"], atol=1e-4)
    assert torch.allclose(poly, c["rb_polygons"], atol=1e-3)
    # + PostProcessorAcademic's text-score filter
    boxes, idx, poly, _ = pp.post_process(c["boxes"], c["scores"], c["text_scores"])
    assert torch.equal(idx, c["idx"])
    assert torch.allclose(boxes, c["out_boxes"], atol=1e-4)
    assert torch.allclose(poly, c["out_polygons"], atol=1e-3)


def test_cases_exercise_the_merge_loop(golden):
    """The fixtures are only worth something if merges actually happen (and more than one round of them)."""
    from oracle import postprocess as pp
    iters = [pp.post_process(c["boxes"], c["scores"], None)[3] for c in golden["cases"]]
    assert max(iters) >= 2 and sum(1 for n in iters if n >= 1) >= 5, iters


def test_empty_and_all_filtered():
    from oracle import postprocess as pp
    b, idx, poly, it = pp.post_process(torch.zeros((0, 5)), torch.zeros((0,)), None)
    assert b.shape == (0, 5) and idx.numel() == 0 and poly.shape == (0, 4, 2) and it == 0
    boxes = torch.tensor([[10., 10., 1.5, 30., 0.], [50., 50., 40., 20., 10.]])
    b, idx, poly, it = pp.post_process(boxes, torch.tensor([0.9, 0.1]), None)
    assert idx.numel() == 0  # first is too thin, second is below the valid score
