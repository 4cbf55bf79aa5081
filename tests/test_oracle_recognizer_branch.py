"""Rows a9-a16 wired together: the oracle's ``recognizer_branch`` (oracle/model.py) against golden vectors written by
the reference's OWN ``MaskRotatedRecognizerHybridHead._forward_recognizer`` driving the reference's own modules, the
recognizer head assembled by the reference's builders from its pretrain config
(tools/make_golden_recognizer_branch.py -> tests/golden/recognizer_branch.pt)."""
import os

import pytest
import torch

from golden_common import force_eos_bias, make_recognizer_branch_inputs, seeded_fill

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "recognizer_branch.pt")
PARTS = ("recognizer_feature_fusion", "hybrid_net", "fusion_net", "recognizer_head")
EOS_EXTRA = 2.2


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


@pytest.mark.parametrize("i", range(4))
def test_recognizer_branch_matches_reference(golden, i):
    from oracle import model as om
    c = golden["cases"][i]
    o = om.GlassOracle()
    for j, name in enumerate(PARTS):
        part = getattr(o.roi_heads, name)
        assert sorted(part.state_dict().keys()) == c["state_keys"][name], f"{name}: parameter names differ from the reference's"
        seeded_fill(part, 500 + 10 * c["seed"] + j)
    if c["eos"]:
        with torch.no_grad():
            force_eos_bias(o.roi_heads.recognizer_head.decoder.recognizer)
            o.roi_heads.recognizer_head.decoder.recognizer.decoder.fc.bias[0] += EOS_EXTRA
    image, p2, p3, boxes = make_recognizer_branch_inputs(c["seed"], c["k"])
    with torch.no_grad():
        probs = o.recognizer_branch(image[None], {"p2": p2, "p3": p3}, boxes)
    if not c["has_text"]:       # zero words: the reference leaves the instances without the field
        assert c["k"] == 0 and tuple(probs.shape) == (0, 26, 97)
        return
    want = c["pred_text_prob"]
    assert probs.shape == want.shape == (c["k"], 26, 97)
    assert torch.equal(probs.argmax(2), want.argmax(2))
    assert torch.allclose(probs, want, rtol=1e-4, atol=1e-6), float((probs - want).abs().max())
    steps = (want.sum(2) > 0).sum(1)
    assert torch.equal((probs.sum(2) > 0).sum(1), steps)
    if c["eos"]:                # the early break fired mid-sequence, for all words of the image at once
        assert 1 < int(steps.max()) < 26 and int(steps.min()) == int(steps.max())
        first = [(want[w].argmax(1) == 0).nonzero()[0].item() for w in range(c["k"])]
        assert len(set(first)) > 1 and max(first) == int(steps.max()) - 1
