"""GPU parity of the rotated RPN (taps T3-T5) against the oracle, stage-wise
teacher-forced: each sub-stage gets the ORACLE's upstream tensors so that discrete decisions
(top-k membership, NMS survivors) are compared on identical inputs."""
import math

import pytest
import torch

from parity_common import close

pytestmark = pytest.mark.gpu


def _oracle_and_feats(seed=0, n=2, h=192, w=256):
    from oracle import model as om
    o = om.build_oracle(seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    feats = {f"p{k}": torch.randn(n, 256, math.ceil(h / 2 ** k), math.ceil(w / 2 ** k), generator=g)
             for k in range(2, 7)}
    # a little structure so the top-k is not a pure noise ranking
    with torch.no_grad():
        for m in [o.proposal_generator.rpn_head.objectness_logits, o.proposal_generator.rpn_head.anchor_deltas]:
            m.weight.mul_(20.0)
    return o, feats, (h, w)


def _oracle_rpn_per_image(o, feats, image_size):
    outs = []
    for i in range(feats["p2"].shape[0]):
        taps = {}
        with torch.no_grad():
            o.rpn({k: v[i:i + 1] for k, v in feats.items()}, image_size, taps)
        outs.append(taps)
    return outs


def _pred_map_from_oracle(taps, feats, A=12, ld=80):
    """Oracle logits [1, HWA] / deltas [1, HWA, 5] per level -> our fused [1,h,w,ld] layout."""
    maps = []
    for lvl, k in enumerate(range(2, 7)):
        h, w = feats[f"p{k}"].shape[-2:]
        m = torch.zeros(1, h, w, ld)
        m[..., :A] = taps["rpn_logits"][lvl].view(1, h, w, A)
        m[..., A:6 * A] = taps["rpn_deltas"][lvl].view(1, h, w, A * 5)
        maps.append(m)
    return maps


def test_rpn_head_matches_oracle(glass_lib):
    from glass_text_spotting_b200 import ops
    from glass_text_spotting_b200.modeling.rpn import B200RotatedRPN
    o, feats, size = _oracle_and_feats()
    taps = _oracle_rpn_per_image(o, feats, size)
    rpn = B200RotatedRPN(o.state_dict())
    preds = rpn.head({k: ops.Act.from_nchw(v.cuda()) for k, v in feats.items()})
    for i in range(2):
        want = _pred_map_from_oracle(taps[i], feats)
        for lvl in range(5):
            close(preds[lvl][i:i + 1, ..., :72], want[lvl][..., :72], f"rpn pred img{i} lvl{lvl}")


def test_rpn_topk_decode_matches_oracle(glass_lib):
    from glass_text_spotting_b200.modeling.rpn import B200RotatedRPN
    o, feats, size = _oracle_and_feats(seed=1)
    taps = _oracle_rpn_per_image(o, feats, size)
    rpn = B200RotatedRPN(o.state_dict())
    for i in range(2):
        maps = [m.cuda().contiguous() for m in _pred_map_from_oracle(taps[i], feats)]
        boxes, scores = rpn.topk_decode(maps)
        want_b, want_s = taps[i]["rpn_topk_boxes"][0], taps[i]["rpn_topk_scores"][0]
        # the oracle concatenates min(1000, HWA) per level; ours pads every level to 1000 slots
        off_o = 0
        for lvl, k in enumerate(range(2, 7)):
            hwa = feats[f"p{k}"].shape[-2] * feats[f"p{k}"].shape[-1] * 12
            cnt = min(1000, hwa)
            got_s = scores[0, lvl * 1000: lvl * 1000 + cnt].cpu()
            got_b = boxes[0, lvl * 1000: lvl * 1000 + cnt].cpu()
            assert torch.equal(got_s, want_s[off_o: off_o + cnt]), f"img{i} lvl{lvl}: top-k scores/order differ"
            close(got_b, want_b[off_o: off_o + cnt], f"img{i} lvl{lvl} topk boxes")
            if cnt < 1000:
                assert torch.isinf(scores[0, lvl * 1000 + cnt: (lvl + 1) * 1000]).all()
            off_o += cnt


def test_rpn_select_matches_oracle(glass_lib):
    from glass_text_spotting_b200.modeling.rpn import B200RotatedRPN
    o, feats, size = _oracle_and_feats(seed=2)
    taps = _oracle_rpn_per_image(o, feats, size)
    rpn = B200RotatedRPN(o.state_dict())
    n = 2
    boxes = torch.zeros(n, 5000, 5)
    scores = torch.full((n, 5000), float("-inf"))
    for i in range(n):
        off_o = 0
        for lvl, k in enumerate(range(2, 7)):
            cnt = min(1000, feats[f"p{k}"].shape[-2] * feats[f"p{k}"].shape[-1] * 12)
            boxes[i, lvl * 1000: lvl * 1000 + cnt] = taps[i]["rpn_topk_boxes"][0, off_o: off_o + cnt]
            scores[i, lvl * 1000: lvl * 1000 + cnt] = taps[i]["rpn_topk_scores"][0, off_o: off_o + cnt]
            off_o += cnt
    hw = torch.tensor([size, size], dtype=torch.float32).cuda()
    ob, os_, oi, oc = rpn.select(boxes.cuda().contiguous(), scores.cuda().contiguous(), hw)
    for i in range(n):
        want_b, want_s = taps[i]["proposal_boxes"], taps[i]["objectness_logits"]
        k = int(oc[i].item())
        assert k == want_b.shape[0], (k, want_b.shape)
        assert torch.equal(os_[i, :k].cpu(), want_s), f"img{i}: kept set / order differs"
        close(ob[i, :k], want_b, f"img{i} proposals")


def test_rpn_end_to_end_from_features(glass_lib):
    """Whole proposal generator from shared features: the kept proposals agree with the oracle's."""
    from glass_text_spotting_b200 import ops
    from glass_text_spotting_b200.modeling.rpn import B200RotatedRPN
    o, feats, size = _oracle_and_feats(seed=3)
    taps = _oracle_rpn_per_image(o, feats, size)
    rpn = B200RotatedRPN(o.state_dict())
    hw = torch.tensor([size, size], dtype=torch.float32).cuda()
    ob, os_, oi, oc = rpn.forward_device({k: ops.Act.from_nchw(v.cuda()) for k, v in feats.items()}, hw)
    for i in range(2):
        want_b = taps[i]["proposal_boxes"]
        k = int(oc[i].item())
        # set agreement: every oracle proposal has a counterpart within tolerance (logit noise ~1e-6 can swap ranks)
        d = (ob[i, :k].cpu()[:, None, :] - want_b[None, :, :]).abs().amax(-1)
        matched = (d.min(0).values < 1e-2).float().mean().item()
        assert matched >= 0.97, f"img{i}: only {matched:.3f} of the oracle proposals reproduced"


# ------------------------------------------------------------------------------ generic rotated NMS
def _rand_boxes(g, n, extent=200.0):
    cx = torch.rand(n, generator=g) * extent
    cy = torch.rand(n, generator=g) * extent
    w = 10 + torch.rand(n, generator=g) * 60
    h = 5 + torch.rand(n, generator=g) * 30
    a = torch.rand(n, generator=g) * 360 - 180
    return torch.stack((cx, cy, w, h, a), 1)


@pytest.mark.parametrize("n,thr", [(100, 0.35), (700, 0.5), (3000, 0.7)])
def test_nms_rotated_matches_oracle(glass_lib, n, thr):
    from glass_text_spotting_b200 import ops
    from oracle import d2_ops
    g = torch.Generator().manual_seed(n)
    boxes, scores = _rand_boxes(g, n), torch.rand(n, generator=g)
    scores[::7] = scores[3]  # ties resolve by index
    keep = d2_ops.nms_rotated(boxes, scores, thr)[:100]
    ob, os_, oi, oc = ops.nms_rotated(boxes[None].cuda().contiguous(), scores[None].cuda().contiguous(), thr, 100)
    k = int(oc[0].item())
    assert k == keep.numel()
    assert torch.equal(oi[0, :k].cpu().long(), keep)
    assert torch.equal(ob[0, :k].cpu(), boxes[keep])


def test_nms_rotated_angle0_equals_torchvision(glass_lib):
    import torchvision
    from glass_text_spotting_b200 import ops
    g = torch.Generator().manual_seed(5)
    b = _rand_boxes(g, 400)
    b[:, 4] = 0
    s = torch.rand(400, generator=g)
    xyxy = torch.stack((b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2), 1)
    keep = torchvision.ops.nms(xyxy, s, 0.5)[:128]
    ob, os_, oi, oc = ops.nms_rotated(b[None].cuda().contiguous(), s[None].cuda().contiguous(), 0.5, 128)
    assert torch.equal(oi[0, :int(oc[0])].cpu().long(), keep)


def test_nms_batched_groups_and_empty(glass_lib):
    from glass_text_spotting_b200 import ops
    from oracle import d2_ops
    g = torch.Generator().manual_seed(6)
    boxes, scores = _rand_boxes(g, 300, extent=80.0), torch.rand(300, generator=g)
    grp = torch.randint(0, 3, (300,), generator=g)
    keep = d2_ops.batched_nms_rotated(boxes, scores, grp, 0.3)[:50]
    ob, os_, oi, oc = ops.nms_rotated(boxes[None].cuda().contiguous(), scores[None].cuda().contiguous(), 0.3, 50,
                                      group=grp[None].int().cuda().contiguous())
    assert torch.equal(oi[0, :int(oc[0])].cpu().long(), keep)
    # all candidates invalid -> count 0
    s2 = torch.full((1, 300), float("-inf")).cuda()
    assert int(ops.nms_rotated(boxes[None].cuda().contiguous(), s2, 0.3, 50)[3][0]) == 0
