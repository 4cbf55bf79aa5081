"""GPU parity of the device post-processor (glass_postprocess_merge / glass_text_scores) against
(a) golden vectors produced by the reference's OWN PostProcessorRotatedBoxes / TextEncoder classes
(tests/golden/postprocess.pt, tools/make_golden_postprocess.py) and (b) the oracle restatement on further seeds.
Kept sets and their order must be identical; boxes agree to 2e-3 px / degrees (cv2.minAreaRect is float32 rotating
calipers, the device takes the exact minimum over point-pair directions in double)."""
import os

import pytest
import torch

from golden_common import make_postprocess_case

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "postprocess.pt")
ATOL = 2e-3


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def _run(ops, boxes, scores, text_scores=None):
    n = boxes.shape[0]
    r = ops.postprocess_merge(boxes.reshape(1, n, 5).cuda().contiguous(), scores.reshape(1, n).cuda().contiguous(),
                              None, None if text_scores is None else text_scores.reshape(1, n).cuda().contiguous())
    k = int(r["count"][0])
    return (r["boxes"][0, :k].cpu(), r["index"][0, :k].cpu().long(), r["polygons"][0, :k].cpu(), int(r["iters"][0]),
            r)


def _same_boxes(got, ref):
    assert got.shape == ref.shape
    if got.numel() == 0:
        return
    d = (got - ref).abs()
    # angles are compared modulo 360 (-180 == 180)
    d[:, 4] = torch.minimum(d[:, 4], (360.0 - d[:, 4]).abs())
    assert d.max().item() <= ATOL, d.max().item()


@pytest.mark.parametrize("i", range(8))
def test_matches_reference_golden(glass_lib, golden, i):
    from glass_text_spotting_b200 import ops
    c = golden["cases"][i]
    boxes, idx, poly, iters, raw = _run(ops, c["boxes"], c["scores"])
    assert torch.equal(idx, c["rb_idx"]), (idx.tolist(), c["rb_idx"].tolist())
    _same_boxes(boxes, c["rb_boxes"])
    assert torch.allclose(poly, c["rb_polygons"], atol=5e-3)
    # padding contract
    k = len(idx)
    assert (raw["index"][0, k:] == -1).all() and (raw["boxes"][0, k:] == 0).all()
    # + PostProcessorAcademic's text filter
    boxes, idx, poly, _, _ = _run(ops, c["boxes"], c["scores"], c["text_scores"])
    assert torch.equal(idx, c["idx"])
    _same_boxes(boxes, c["out_boxes"])


@pytest.mark.parametrize("seed", [21, 22, 23, 24, 25, 26])
def test_matches_oracle_on_more_seeds(glass_lib, seed):
    from glass_text_spotting_b200 import ops
    from oracle import postprocess as opp
    boxes, scores = make_postprocess_case(seed, 10 + seed % 13, 15 + seed % 7)
    boxes, scores = boxes[:100], scores[:100]
    ref_boxes, ref_idx, ref_poly, ref_iters = opp.post_process(boxes, scores, None)
    got_boxes, got_idx, got_poly, got_iters, _ = _run(ops, boxes, scores)
    assert torch.equal(got_idx, ref_idx)
    assert got_iters == ref_iters
    _same_boxes(got_boxes, ref_boxes)
    assert torch.allclose(got_poly, ref_poly, atol=5e-3)


def test_batched_with_counts(glass_lib, golden):
    """All golden cases in one launch (one CTA per image), ragged through ``counts``; rows past a count are ignored."""
    from glass_text_spotting_b200 import ops
    cases = golden["cases"]
    m = 100
    boxes = torch.full((len(cases), m, 5), 7.0)   # junk beyond the counts must not matter
    scores = torch.full((len(cases), m), 0.99)
    ts = torch.full((len(cases), m), 0.99)
    counts = torch.zeros(len(cases), dtype=torch.int32)
    for i, c in enumerate(cases):
        n = len(c["boxes"])
        boxes[i, :n], scores[i, :n], ts[i, :n], counts[i] = c["boxes"], c["scores"], c["text_scores"], n
    r = ops.postprocess_merge(boxes.cuda(), scores.cuda(), counts.cuda(), ts.cuda())
    for i, c in enumerate(cases):
        k = int(r["count"][i])
        assert torch.equal(r["index"][i, :k].cpu().long(), c["idx"]), i
        _same_boxes(r["boxes"][i, :k].cpu(), c["out_boxes"])


def test_text_scores_match_reference(glass_lib, golden):
    """glass_text_scores against the word scores the reference's TextEncoder.decode_prod_v2 produced."""
    from glass_text_spotting_b200 import ops
    for c in golden["cases"]:
        pmax, pidx = c["text_prob_max"], c["text_prob_idx"]
        n, steps = pmax.shape
        probs = torch.zeros(n, steps, 97)
        probs.scatter_(2, pidx.unsqueeze(2), pmax.unsqueeze(2))
        score, idx, maxp = ops.text_scores(probs.cuda().contiguous(), 1, want_steps=True)
        assert torch.equal(idx.cpu().long(), pidx) and torch.equal(maxp.cpu(), pmax)
        assert torch.allclose(score.cpu(), c["text_scores"], rtol=1e-6, atol=1e-12)
    # rows zeroed by the decoder's early break (all-zero steps: argmax 0, prob 0) and words without a stop symbol
    g = torch.Generator().manual_seed(5)
    probs = torch.softmax(torch.randn(6, 26, 97, generator=g) * 4, dim=2)
    probs[:, :, 1] = 0.0       # no stop symbol anywhere
    probs[0, 10:] = 0.0        # early break rows
    from glass_text_spotting_b200.text import TextDecoder
    ref = TextDecoder().decode_probs(probs)
    score = ops.text_scores(probs.cuda().contiguous(), 1).cpu()
    assert torch.allclose(score, torch.tensor([w["score"] for w in ref], dtype=torch.float32), rtol=1e-6, atol=1e-30)


def test_instances_api_and_edge_cases(glass_lib, golden):
    from glass_text_spotting_b200.postprocess import B200PostProcessor, PostProcessingConfig
    from glass_text_spotting_b200.structures import Instances, RotatedBoxes
    c = golden["cases"][1]
    n = len(c["boxes"])
    probs = torch.zeros(n, 26, 97)
    probs.scatter_(2, c["text_prob_idx"].unsqueeze(2), c["text_prob_max"].unsqueeze(2))
    inst = Instances((1024, 1024), pred_boxes=RotatedBoxes(c["boxes"].cuda()), scores=c["scores"].cuda(),
                     pred_classes=torch.zeros(n, dtype=torch.int64).cuda(), pred_text_prob=probs.cuda(),
                     orig=torch.arange(n).cuda())
    out = B200PostProcessor()(inst)
    assert torch.equal(out.orig.cpu(), c["idx"])
    assert out.pred_polygons.shape == (len(c["idx"]), 4, 2) and out.pred_text_prob.shape[0] == len(c["idx"])
    _same_boxes(out.pred_boxes.tensor.cpu(), c["out_boxes"])
    # skip-all, no text filter, empty input
    assert B200PostProcessor(PostProcessingConfig(SKIP_ALL=True))(inst) is inst
    out2 = B200PostProcessor(text_filter=False)(inst)
    assert torch.equal(out2.orig.cpu(), c["rb_idx"])
    empty = inst[torch.zeros(0, dtype=torch.int64).cuda()]
    oe = B200PostProcessor()(empty)
    assert len(oe) == 0 and oe.pred_polygons.shape == (0, 4, 2)
    with pytest.raises(AssertionError):
        B200PostProcessor(PostProcessingConfig(VALID_CONFIDENCE=0.5, DETECT_THRESHOLD=0.25))
